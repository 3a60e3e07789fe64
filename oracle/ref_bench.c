/* ref_bench.c -- CPU arm of bench.py (TEST / MEASUREMENT INFRASTRUCTURE): the reference's per-candidate hot loop
 *   InterSearch::xMVReprojectionInterpolation (InterSearch.cpp:6277-6385) + m_cDistParam.distFunc (InterSearch.cpp:363-453)
 * timed on the host cores, multi-threaded over blocks.  Interpolation and distortion run the REFERENCE'S OWN compiled code
 * (oracle/_ref/libvtmref.so: InterpolationFilter / RdCost with their AVX2 units, called per 4x4 through filterHor/filterVer
 * and once per candidate through setDistParam + distFunc, exactly as the reference does); the float mapping
 * (reprojectMotionVectorSubblocks) is the C restatement of oracle/mm_oracle.c, because the reference's mapping needs
 * Eigen 3.3.7, which is not on the box (the stand-in used for the parity builds evaluates expressions far too slowly to be
 * a fair timing).  Built into oracle/_ref/librefbench.so by oracle/Makefile where /root/reference exists. */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mm_oracle.h"

void ref_init(int level);
void ref_filter_hor(int compID, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h, int frac, int isLast,
                    int bd, int lo, int hi);
void ref_filter_ver(int compID, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h, int frac, int isFirst,
                    int isLast, int bd, int lo, int hi);
unsigned long long ref_dist(const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h, int bd, int subShiftMode,
                            int useHadamard);

typedef struct { int32_t x, y, w, h, model, ref_slot, rot_slot, cand_first, cand_count, org_x, org_y, reserved; } rb_block;
typedef struct { int32_t hor, ver; } rb_mv;

typedef struct
{
  OrcCfg cfg;
  const float* R;           /* GEODESIC_CAMPOSE rotation */
  const int16_t* ref00; int ref_stride;
  const int16_t* org; int org_stride;
  const rb_block* blocks; int n_blocks;
  const rb_mv* mvs; int mv_shift;
  int use_had;
  double budget_s;
  volatile long next;       /* shared block counter (wraps around the batch) */
  uint64_t* dist;           /* optional: distortions of the first pass over the batch */
} rb_job;

typedef struct { rb_job* job; double samples, t_map, t_interp, t_dist; long blocks; } rb_thread;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* one candidate: InterSearch.cpp:6277-6385 */
static uint64_t eval_candidate(OrcCtx* c, const rb_job* J, const rb_block* b, int mvh, int mvv, int32_t* px, int32_t* py, int16_t* tmp,
                               int16_t* tmp2, double* tm, double* ti, double* td)
{
  const double t0 = now();
  orc_reproject(c, b->x, b->y, b->w, b->h, mvh, mvv, b->model, 0, 0, px, py);
  const double t1 = now();
  const int rows = b->h / 4, cols = b->w / 4, W = J->cfg.pic_w, H = J->cfg.pic_h, rs = J->ref_stride;
  for (int col = 0; col < cols; ++col)
    for (int row = 0; row < rows; ++row)
    {
      const int i = col * rows + row;
      const int xP = px[i] >> 4, yP = py[i] >> 4, xF = px[i] & 15, yF = py[i] & 15;
      int16_t* d = tmp + row * 4 * 128 + col * 4;
      if (xP < 0 || yP < 0 || xP >= W - 4 || yP >= H - 4)
      {
        for (int y = 0; y < 4; ++y) memset(d + y * 128, 0, 8);
        continue;
      }
      const int16_t* p = J->ref00 + (long)yP * rs + xP;
      if (yF == 0) ref_filter_hor(0, p, rs, d, 128, 4, 4, xF, 1, 10, 0, 1023);
      else if (xF == 0) ref_filter_ver(0, p, rs, d, 128, 4, 4, yF, 1, 1, 10, 0, 1023);
      else
      {
        ref_filter_hor(0, p - 3 * rs, rs, tmp2, 4, 4, 11, xF, 0, 10, 0, 1023);
        ref_filter_ver(0, tmp2 + 3 * 4, 4, d, 128, 4, 4, yF, 0, 1, 10, 0, 1023);
      }
    }
  const double t2 = now();
  const uint64_t dist = ref_dist(J->org + (long)b->org_y * J->org_stride + b->org_x, J->org_stride, tmp, 128, b->w, b->h, 10, 0, J->use_had);
  const double t3 = now();
  *tm += t1 - t0; *ti += t2 - t1; *td += t3 - t2;
  return dist;
}

static void* worker(void* arg)
{
  rb_thread* T = (rb_thread*)arg;
  rb_job* J = T->job;
  OrcCtx* c = orc_create(&J->cfg);
  if (J->R) orc_set_rotation(c, 10, J->R);
  int32_t* px = (int32_t*)malloc(sizeof(int32_t) * 2048);
  int32_t* py = px + 1024;
  int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * 128 * 128);
  int16_t tmp2[4 * 11];
  const double t0 = now();
  while (now() - t0 < J->budget_s)
  {
    const long k = __sync_fetch_and_add(&J->next, 1);
    const rb_block* b = &J->blocks[k % J->n_blocks];
    for (int i = 0; i < b->cand_count; ++i)
    {
      const rb_mv mv = J->mvs[b->cand_first + i];
      const uint64_t d = eval_candidate(c, J, b, mv.hor * (1 << J->mv_shift), mv.ver * (1 << J->mv_shift), px, py, tmp, tmp2, &T->t_map,
                                        &T->t_interp, &T->t_dist);
      if (J->dist && k < J->n_blocks) J->dist[b->cand_first + i] = d;
    }
    T->samples += (double)b->w * b->h * b->cand_count;
    T->blocks++;
  }
  free(px); free(tmp);
  orc_destroy(c);
  return 0;
}

/* returns luma sample-candidates per second; stats[0..4] = seconds, blocks, t_map, t_interp, t_dist (thread-seconds) */
double refbench_run(const OrcCfg* cfg, const float* R, const int16_t* ref00, int ref_stride, const int16_t* org, int org_stride,
                    const rb_block* blocks, int n_blocks, const rb_mv* mvs, int mv_shift, int use_had, int n_threads, double budget_s,
                    uint64_t* dist_out, double* stats)
{
  ref_init(-1);
  rb_job J;
  memset(&J, 0, sizeof(J));
  J.cfg = *cfg; J.R = R; J.ref00 = ref00; J.ref_stride = ref_stride; J.org = org; J.org_stride = org_stride; J.blocks = blocks;
  J.n_blocks = n_blocks; J.mvs = mvs; J.mv_shift = mv_shift; J.use_had = use_had; J.budget_s = budget_s; J.dist = dist_out;
  rb_thread* T = (rb_thread*)calloc((size_t)n_threads, sizeof(rb_thread));
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  const double t0 = now();
  for (int i = 0; i < n_threads; ++i) { T[i].job = &J; pthread_create(&th[i], 0, worker, &T[i]); }
  for (int i = 0; i < n_threads; ++i) pthread_join(th[i], 0);
  const double dt = now() - t0;
  double samples = 0, blocks_done = 0, tm = 0, ti = 0, td = 0;
  for (int i = 0; i < n_threads; ++i) { samples += T[i].samples; blocks_done += T[i].blocks; tm += T[i].t_map; ti += T[i].t_interp; td += T[i].t_dist; }
  if (stats) { stats[0] = dt; stats[1] = blocks_done; stats[2] = tm; stats[3] = ti; stats[4] = td; }
  free(T); free(th);
  return samples / dt;
}

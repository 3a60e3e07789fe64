/* mm_oracle.c -- CPU oracle (plain C restatement of the reference).  See mm_oracle.h for the
 * rules: TEST INFRASTRUCTURE ONLY; float mapping is "parity unpinned" where it depends on
 * Eigen 3.3.7 (not on the box), libm results come from the real glibc of the box.
 *
 * Build: gcc -O2 -msse4.1 -ffp-contract=off -fno-builtin (oracle/Makefile) -- the reference is
 * built -O3 -msse4.1 without fast-math (CMakeLists.txt:92), i.e. no FMA contraction anywhere.
 */
#include "mm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_F ((float)M_PI)               /* TCoord(M_PI) */
#define TWO_PI_F ((float)2 * (float)M_PI) /* TCoord(2)*TCoord(M_PI) */
#define PI_2_F ((float)M_PI_2)

/* ---------------------------------------------------------------------------------------------
 * Eigen 3.3.7 Core/arch/SSE/MathFunctions.h psin/pcos<Packet4f>, one lane.
 * Used for Array<float>::sin()/cos() on linear indices < (n/4)*4 of a fully vectorisable
 * expression (LinearVectorizedTraversal, aligned destination).  UNPINNED (Eigen not on box).
 * ------------------------------------------------------------------------------------------- */
static int32_t cvtt_ps(float x) /* _mm_cvttps_epi32 lane: 0x80000000 when out of range / NaN */
{
  if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int32_t)0x80000000u;
  return (int32_t)x;
}
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static void eig_ps_core(float ax, int32_t* jo, float* ys_o, float* yc_o)
{
  const float FOPI = 1.27323954473516f;
  const float mDP1 = -0.78515625f, mDP2 = -2.4187564849853515625e-4f, mDP3 = -3.77489497744594108e-8f;
  float y = ax * FOPI;
  int32_t j = cvtt_ps(y);
  j = (int32_t)(((uint32_t)j + 1u) & ~1u);
  y = (float)j;
  float x = ax;
  float t1 = y * mDP1, t2 = y * mDP2, t3 = y * mDP3;
  x = x + t1; x = x + t2; x = x + t3;
  float z = x * x;
  float yc = 2.443315711809948E-005f;
  yc = yc * z + -1.388731625493765E-003f;
  yc = yc * z + 4.166664568298827E-002f;
  yc = yc * z; yc = yc * z;
  float tmp = z * 0.5f;
  yc = yc - tmp; yc = yc + 1.0f;
  float ys = -1.9515295891E-4f;
  ys = ys * z + 8.3321608736E-3f;
  ys = ys * z + -1.6666654611E-1f;
  ys = ys * z; ys = ys * x; ys = ys + x;
  *jo = j; *ys_o = ys; *yc_o = yc;
}
static float eig_psin(float x)
{
  uint32_t sign = f2u(x) & 0x80000000u;
  int32_t j; float ys, yc;
  eig_ps_core(fabsf(x), &j, &ys, &yc);
  uint32_t swap = ((uint32_t)j & 4u) << 29;
  float r = (((uint32_t)j & 2u) == 0u) ? ys : yc;
  return u2f(f2u(r) ^ (sign ^ swap));
}
static float eig_pcos(float x)
{
  int32_t j; float ys, yc;
  eig_ps_core(fabsf(x), &j, &ys, &yc);
  uint32_t j2 = (uint32_t)j - 2u;
  uint32_t sign = ((~j2) & 4u) << 29;
  float r = ((j2 & 2u) == 0u) ? ys : yc;
  return u2f(f2u(r) ^ sign);
}
/* Array sin/cos inside a vectorisable expression: pk = element lies in the packet body. */
static float a_sin(float x, int pk) { return pk ? eig_psin(x) : sinf(x); }
static float a_cos(float x, int pk) { return pk ? eig_pcos(x) : cosf(x); }

/* Eigen scalar cwiseMin/cwiseMax = std::min/std::max (NaN in first operand propagates), R7. */
static float eig_min(float a, float b) { return (b < a) ? b : a; }
static float eig_max(float a, float b) { return (a < b) ? b : a; }

/* 3-term dot product as Eigen's unrolled redux: a0*b0 + (a1*b1 + a2*b2). */
static float dot3(float a0, float b0, float a1, float b1, float a2, float b2)
{
  float p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2;
  float s12 = p1 + p2;
  return p0 + s12;
}
static void mat3_mul(const float A[9], const float B[9], float C[9]) /* row-major, lazy coeff product */
{
  float T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = dot3(A[i * 3 + 0], B[0 * 3 + j], A[i * 3 + 1], B[1 * 3 + j], A[i * 3 + 2], B[2 * 3 + j]);
  memcpy(C, T, sizeof(T));
}
static void mat3_vec(const float A[9], float x, float y, float z, float o[3])
{
  o[0] = dot3(A[0], x, A[1], y, A[2], z);
  o[1] = dot3(A[3], x, A[4], y, A[5], z);
  o[2] = dot3(A[6], x, A[7], y, A[8], z);
}
static void mat3_tvec(const float A[9], float x, float y, float z, float o[3]) /* A^T * v */
{
  o[0] = dot3(A[0], x, A[3], y, A[6], z);
  o[1] = dot3(A[1], x, A[4], y, A[7], z);
  o[2] = dot3(A[2], x, A[5], y, A[8], z);
}

/* Eigen::AngleAxis<float>::toRotationMatrix (Eigen/src/Geometry/AngleAxis.h), row-major out. */
static void angle_axis(float angle, float ax, float ay, float az, float R[9])
{
  float s = sinf(angle), c = cosf(angle);
  float sx = s * ax, sy = s * ay, sz = s * az;
  float omc = 1.0f - c;
  float cx = omc * ax, cy = omc * ay, cz = omc * az;
  float tmp;
  tmp = cx * ay; R[0 * 3 + 1] = tmp - sz; R[1 * 3 + 0] = tmp + sz;
  tmp = cx * az; R[0 * 3 + 2] = tmp + sy; R[2 * 3 + 0] = tmp - sy;
  tmp = cy * az; R[1 * 3 + 2] = tmp - sx; R[2 * 3 + 1] = tmp + sx;
  R[0] = cx * ax + c; R[4] = cy * ay + c; R[8] = cz * az + c;
}

/* ---------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------- */
struct OrcCtx {
  OrcCfg cfg;
  float focal;      /* EquirectangularProjection: TCoord(1./tan(M_PI/H)), Projection.h:128 */
  float ang_res;    /* TCoord(M_PI / H), MVReprojection.cpp:27,33,39 */
  int rows4, cols4; /* H/4, W/4 */
  float* mpa_x[4];  /* [model 1..3] column-major perspective caches */
  float* mpa_y[4];
  uint8_t* mpa_vip[4];
  float R[ORC_NUM_MODELS][9]; /* GeodesicMotionModel::m_rotationMatrix per GED model */
};

/* CoordinateConversion::cartesianToSpherical, array version (Coordinate.cpp:36-44). */
static void arr_cart_to_sph(float x, float y, float z, float* r, float* theta, float* phi)
{
  float rr = sqrtf((x * x + y * y) + z * z);
  *r = rr;
  *theta = acosf(eig_max(eig_min(z / rr, 1.0f), -1.0f));
  *phi = atan2f(y, x);
}
/* scalar version (Coordinate.cpp:46-51): std::min/std::max argument order differs -> NaN is clamped. */
static float std_minf(float a, float b) { return (b < a) ? b : a; }
static float std_maxf(float a, float b) { return (a < b) ? b : a; }
static void sc_cart_to_sph(const float c[3], float s[3])
{
  float r = sqrtf((c[0] * c[0]) + (c[1] * c[1]) + (c[2] * c[2]));
  s[0] = r;
  s[1] = acosf(std_minf(1.0f, std_maxf(-1.0f, c[2] / r)));
  s[2] = atan2f(c[1], c[0]);
}
/* CoordinateConversion::sphericalToCartesian, array (Coordinate.cpp:53-60) and scalar (:62-68). */
static void arr_sph_to_cart(float r, float theta, float phi, int pk, float o[3])
{
  o[0] = (r * a_sin(theta, pk)) * a_cos(phi, pk);
  o[1] = (r * a_sin(theta, pk)) * a_sin(phi, pk);
  o[2] = r * a_cos(theta, pk);
}
static void sc_sph_to_cart(float r, float theta, float phi, float o[3])
{
  o[0] = r * sinf(theta) * cosf(phi);
  o[1] = r * sinf(theta) * sinf(phi);
  o[2] = r * cosf(theta);
}

/* EquirectangularProjection::toSphere array (Projection.cpp:213-221) / scalar (:223-228), pixelOffset 0. */
static void erp_to_sphere_arr(const OrcCtx* c, float X, float Y, int pk, float o[3])
{
  float x = X + 0.0f, y = Y + 0.0f;
  float phi = ((-(x / (float)c->cfg.pic_w)) * (float)2) * PI_F;
  float theta = (y / (float)c->cfg.pic_h) * PI_F;
  arr_sph_to_cart(1.0f, theta, phi, pk, o);
}
static void erp_to_sphere_sc(const OrcCtx* c, float X, float Y, float o[3])
{
  float phi = -((X + 0.0f) / (float)c->cfg.pic_w) * (float)2 * PI_F;
  float theta = ((Y + 0.0f) / (float)c->cfg.pic_h) * PI_F;
  sc_sph_to_cart(1.0f, theta, phi, o);
}
/* EquirectangularProjection::fromSphere array (Projection.cpp:230-240). */
static void erp_from_sphere_arr(const OrcCtx* c, float x, float y, float z, float* X, float* Y)
{
  float r, theta, phi;
  arr_cart_to_sph(x, y, z, &r, &theta, &phi);
  if (phi > 0) phi = phi - TWO_PI_F;
  *X = (-(phi / TWO_PI_F)) * (float)c->cfg.pic_w - 0.0f;
  *Y = (theta / PI_F) * (float)c->cfg.pic_h - 0.0f;
}

/* PerspectiveProjection::fromSphere array (Projection.cpp:158-178), optical centre (0,0). */
static void pers_from_sphere_arr(const OrcCtx* c, float sx, float sy, float sz, int pk, float* X, float* Y, uint8_t* vip)
{
  float r, theta, phi;
  arr_cart_to_sph(sy, -sz, -sx, &r, &theta, &phi);
  float polarR = c->focal * tanf(theta);
  *X = polarR * a_cos(phi, pk) + 0.0f;
  *Y = polarR * a_sin(phi, pk) + 0.0f;
  *vip = polarR < 0;
}
/* PerspectiveProjection::toSphere array (Projection.cpp:119-142). */
static void pers_to_sphere_arr(const OrcCtx* c, float X, float Y, int vip, int pk, float o[3])
{
  float x = X - 0.0f, y = Y - 0.0f;
  float polarR = sqrtf(x * x + y * y);
  float polarPhi = atan2f(y, x);
  float theta = atanf(polarR / c->focal);
  float vf = vip ? 1.0f : 0.0f;
  theta = theta - vf * ((float)2. * theta - PI_F);
  float phi = polarPhi - vf * PI_F;
  float s[3];
  arr_sph_to_cart(1.0f, theta, phi, pk, s);
  o[0] = -s[2]; o[1] = s[0]; o[2] = -s[1];
}

/* MotionPlaneAdaptiveMotionModel::toPerspective (MotionPlaneAdaptiveMotionModel.cpp:106-135). */
static void mpa_to_perspective(const OrcCtx* c, int model, float X, float Y, int pk, float* PX, float* PY, uint8_t* vip)
{
  float s[3];
  erp_to_sphere_arr(c, X, Y, pk, s);
  float mx, my, mz;
  if (model == ORC_MPA_FRONT_BACK) { mx = s[0]; my = s[1]; mz = s[2]; }
  else if (model == ORC_MPA_LEFT_RIGHT) { mx = s[1]; my = -s[0]; mz = s[2]; }
  else { mx = -s[2]; my = s[1]; mz = s[0]; }
  pers_from_sphere_arr(c, mx, my, mz, pk, PX, PY, vip);
}
/* MotionPlaneAdaptiveMotionModel::toProjection (:163-187). */
static void mpa_to_projection(const OrcCtx* c, int model, float PX, float PY, int vip, int pk, float* X, float* Y)
{
  float m[3], sx, sy, sz;
  pers_to_sphere_arr(c, PX, PY, vip, pk, m);
  if (model == ORC_MPA_FRONT_BACK) { sx = m[0]; sy = m[1]; sz = m[2]; }
  else if (model == ORC_MPA_LEFT_RIGHT) { sx = -m[1]; sy = m[0]; sz = m[2]; }
  else { sx = m[2]; sy = m[1]; sz = -m[0]; }
  erp_from_sphere_arr(c, sx, sy, sz, X, Y);
}

OrcCtx* orc_create(const OrcCfg* cfg)
{
  OrcCtx* c = (OrcCtx*)calloc(1, sizeof(OrcCtx));
  c->cfg = *cfg;
  c->focal = (float)(1. / tan(M_PI / cfg->pic_h));
  c->ang_res = (float)(M_PI / cfg->pic_h);
  c->rows4 = cfg->pic_h / 4;
  c->cols4 = cfg->pic_w / 4;
  const long n = (long)c->rows4 * c->cols4;
  /* MPA caches: toPerspective over the picture-wide coordinate arrays (MVReprojection.cpp:20-24,
   * 67-71); column-major, n is a multiple of 4 for every legal picture size -> packet body. */
  for (int m = ORC_MPA_FRONT_BACK; m <= ORC_MPA_TOP_BOTTOM; ++m)
  {
    c->mpa_x[m] = (float*)malloc(n * sizeof(float));
    c->mpa_y[m] = (float*)malloc(n * sizeof(float));
    c->mpa_vip[m] = (uint8_t*)malloc(n);
    const long body = n & ~3L;
    for (int col = 0; col < c->cols4; ++col)
      for (int row = 0; row < c->rows4; ++row)
      {
        long i = (long)col * c->rows4 + row;
        float X = (float)(4 * col) + cfg->offset4x4, Y = (float)(4 * row) + cfg->offset4x4;
        mpa_to_perspective(c, m, X, Y, i < body, &c->mpa_x[m][i], &c->mpa_y[m][i], &c->mpa_vip[m][i]);
      }
  }
  /* fixed GEDA axes, MVReprojection.cpp:42-52 */
  const float ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1}, e0[3] = {0, 0, 0};
  orc_rodrigues(ex, c->R[ORC_GEODESIC_X]);
  orc_rodrigues(ey, c->R[ORC_GEODESIC_Y]);
  orc_rodrigues(ez, c->R[ORC_GEODESIC_Z]);
  orc_rodrigues(e0, c->R[ORC_GEODESIC_CAMPOSE]); /* default list entry EpipoleList.h:15-17 -> NaN matrix */
  return c;
}
void orc_destroy(OrcCtx* c)
{
  if (!c) return;
  for (int m = 1; m <= 3; ++m) { free(c->mpa_x[m]); free(c->mpa_y[m]); free(c->mpa_vip[m]); }
  free(c);
}
const float* orc_mpa_cache_x(OrcCtx* c, int model) { return c->mpa_x[model]; }
const float* orc_mpa_cache_y(OrcCtx* c, int model) { return c->mpa_y[model]; }
const uint8_t* orc_mpa_cache_vip(OrcCtx* c, int model) { return c->mpa_vip[model]; }

float orc_fixed_to_float(int fixed, int precision)
{
  return (float)(fixed >> precision) + (float)(fixed & ((1 << precision) - 1)) / (float)(1 << precision);
}

void orc_rodrigues(const float e[3], float R[9])
{
  /* GeodesicMotionModel.cpp:25-45.  polarAxisNormalized is a lazy `epipole / norm` expression. */
  const float norm = sqrtf((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]));
  const float px = e[0] / norm, py = e[1] / norm, pz = e[2] / norm;
  const float cx = -py, cy = px, cz = 0.0f;
  const float s = sqrtf(cx * cx + cy * cy + cz * cz);
  float M[9];
  if (s == 0)
  {
    const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(M, I, sizeof(I));
    if (pz < 0) M[8] = -1;
    memcpy(R, M, sizeof(M));
    return;
  }
  const float cc = std_maxf(-1.0f, std_minf(1.0f, pz));
  float K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  K[0 * 3 + 1] = -cz; K[0 * 3 + 2] = cy;
  K[1 * 3 + 0] = cz;  K[1 * 3 + 2] = -cx;
  K[2 * 3 + 0] = -cy; K[2 * 3 + 1] = cx;
  float KK[9];
  mat3_mul(K, K, KK);
  const float f = (1 - cc) / (s * s);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      float id = (i == j) ? 1.0f : 0.0f;
      M[i * 3 + j] = (id + K[i * 3 + j]) + KK[i * 3 + j] * f;
    }
  /* transposeInPlace */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = M[j * 3 + i];
}
void orc_set_rotation(OrcCtx* c, int model, const float R[9]) { memcpy(c->R[model], R, 9 * sizeof(float)); }

/* GeodesicMotionModel::modelGeodesicMotion (GeodesicMotionModel.cpp:101-126): per-call scalar k. */
static float ged_k(const OrcCtx* c, const float* R, float mvx, float bcx, float bcy)
{
  float cc[3], cr[3], sc[3];
  erp_to_sphere_sc(c, bcx, bcy, cc);
  mat3_vec(R, cc[0], cc[1], cc[2], cr);
  sc_cart_to_sph(cr, sc);
  return sinf(sc[1] + c->ang_res * mvx) / sinf(c->ang_res * mvx);
}

/* ---------------------------------------------------------------------------------------------
 * MVReprojection::reprojectMotionVectorSubblocks (MVReprojection.cpp:80-166)
 * ------------------------------------------------------------------------------------------- */
static int scale_x(int comp, int fmt) { return (comp == 0) ? 0 : ((fmt == ORC_CHROMA_420 || fmt == ORC_CHROMA_422) ? 1 : 0); }
static int scale_y(int comp, int fmt) { return (comp == 0) ? 0 : ((fmt == ORC_CHROMA_420) ? 1 : 0); }

void orc_reproject(OrcCtx* c, int pos_x, int pos_y, int size_w, int size_h, int mv_hor, int mv_ver, int model,
                   int comp, int chroma_fmt, int32_t* out_x, int32_t* out_y)
{
  const int csx = scale_x(comp, chroma_fmt), csy = scale_y(comp, chroma_fmt);
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const float scaleX = (float)(1 << csx), scaleY = (float)(1 << csy); /* std::pow(2.f, scale) */
  const int shiftHor = 4 + csx, shiftVer = 4 + csy;
  const int rows = size_h / sbh, cols = size_w / sbw;
  const int n = rows * cols, body = n & ~3;
  const float off = c->cfg.offset4x4;
  const int is_luma = (comp == 0);

  /* :123-124 motion vector as floating point */
  const float mvX = (float)(mv_hor >> 4) + (float)(mv_hor & 15) / (float)16;
  const float mvY = (float)(mv_ver >> 4) + (float)(mv_ver & 15) / (float)16;
  /* :133 block centre, in units of blocks[compID] (quirk Q1) */
  const float bcx = (float)pos_x + ((float)size_w - 1) / (float)2;
  const float bcy = (float)pos_y + ((float)size_h - 1) / (float)2;
  const float dA = c->ang_res;
  const int mv_zero = (mvX == 0 && mvY == 0);
  const int is_mpa = (model >= ORC_MPA_FRONT_BACK && model <= ORC_MPA_TOP_BOTTOM);
  const int is_ged = (model >= ORC_GEODESIC_X && model <= ORC_GEODESIC_CAMPOSE);
  const float* R = c->R[model < ORC_NUM_MODELS ? model : 0];

  /* per-call scalars */
  float d3[3] = {0, 0, 0};        /* 3DT motion vector in 3-D */
  float Rrot[9];                  /* ROT rotationMatrixReally */
  float epsC = 0, alphaC = 0, sinC = 0, cosC = 0; /* TAN */
  float k = 0;                    /* GED modulated */
  if (model == ORC_THREE_D_TRANSLATIONAL && !mv_zero)
  {
    /* ThreeDTranslationalMotionModel.cpp:13-16 */
    float a[3], b[3];
    erp_to_sphere_sc(c, bcx, bcy, a);
    erp_to_sphere_sc(c, bcx + mvX, bcy + mvY, b);
    d3[0] = b[0] - a[0]; d3[1] = b[1] - a[1]; d3[2] = b[2] - a[2];
  }
  else if (model == ORC_ROTATIONAL && !mv_zero)
  {
    /* RotationalMotionModel.cpp:34-52 */
    float cc[3], sc[3], Rz[9], Ry[9], rot[9], uphi[9], uth[9], unrot[9], unrotT[9], t[9];
    erp_to_sphere_sc(c, bcx, bcy, cc);
    sc_cart_to_sph(cc, sc);
    angle_axis(-mvX * dA, 0, 0, 1, Rz);
    angle_axis(mvY * dA, 0, 1, 0, Ry);
    mat3_mul(Rz, Ry, rot);
    angle_axis(-sc[2], 0, 0, 1, uphi);
    angle_axis((float)(M_PI_2 - sc[1]), 0, 1, 0, uth);
    mat3_mul(uth, uphi, unrot);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) unrotT[i * 3 + j] = unrot[j * 3 + i];
    mat3_mul(rot, unrot, t);
    mat3_mul(unrotT, t, Rrot);
  }
  else if (model == ORC_TANGENTIAL && !mv_zero)
  {
    /* TangentialMotionModel.cpp:13-17; unqualified sin/cos bind to the double overloads (R4) */
    float cc[3], sc[3];
    erp_to_sphere_sc(c, bcx, bcy, cc);
    sc_cart_to_sph(cc, sc);
    epsC = (float)(M_PI_2 - sc[1]);
    alphaC = sc[2];
    sinC = (float)sin((double)epsC);
    cosC = (float)cos((double)epsC);
  }
  else if (is_ged && c->cfg.ged_flavor == ORC_GED_MODULATED)
  {
    if (is_luma || !mv_zero) k = ged_k(c, R, mvX, bcx, bcy);
  }

  for (int col = 0; col < cols; ++col)
    for (int row = 0; row < rows; ++row)
    {
      const int i = col * rows + row;
      const int pk = i < body;
      /* :101-120 coordinates of the sub-block (luma: picture cache block; chroma: LinSpaced, step 4) */
      float X, Y;
      if (is_luma) { X = (float)(4 * (pos_x / 4 + col)) + off; Y = (float)(4 * (pos_y / 4 + row)) + off; }
      else
      {
        const float sx0 = (float)pos_x * scaleX + off, sy0 = (float)pos_y * scaleY + off;
        const float sx1 = sx0 + (float)(size_w - sbw) * scaleX, sy1 = sy0 + (float)(size_h - sbh) * scaleY;
        /* Eigen linspaced_op: low + i*step, last element = high; step = (high-low)/(n-1) */
        const float stx = (cols == 1) ? 0.0f : (sx1 - sx0) / (float)(cols - 1);
        const float sty = (rows == 1) ? 0.0f : (sy1 - sy0) / (float)(rows - 1);
        X = (cols > 1 && col == cols - 1) ? sx1 : sx0 + (float)col * stx;
        Y = (rows > 1 && row == rows - 1) ? sy1 : sy0 + (float)row * sty;
      }
      float MX = X, MY = Y;
      if (is_mpa)
      {
        float PX, PY; uint8_t vip;
        if (is_luma)
        { /* modelMotionCached, MotionPlaneAdaptiveMotionModel.cpp:39-69 */
          const long ci = (long)(pos_x / 4 + col) * c->rows4 + (pos_y / 4 + row);
          PX = c->mpa_x[model][ci]; PY = c->mpa_y[model][ci]; vip = c->mpa_vip[model][ci];
        }
        else mpa_to_perspective(c, model, X, Y, pk, &PX, &PY, &vip); /* modelMotion :26-37 */
        const float sgn = vip ? -1.0f : 1.0f;
        const float QX = PX + mvX * sgn, QY = PY + mvY * sgn;
        mpa_to_projection(c, model, QX, QY, vip, pk, &MX, &MY);
      }
      else if (is_ged)
      {
        if (is_luma || !mv_zero) /* un-cached modelMotion returns cart2D for mv==0 (:130-132) */
        {
          float s[3], rp[3], r, th, ph;
          erp_to_sphere_arr(c, X, Y, pk, s);                       /* toRotatedSphere :48-75 */
          mat3_vec(R, s[0], s[1], s[2], rp);
          arr_cart_to_sph(rp[0], rp[1], rp[2], &r, &th, &ph);
          float thM;
          if (c->cfg.ged_flavor == ORC_GED_ORIGINAL) thM = th + dA * mvX;
          else thM = th + atanf(sinf(th) / (k - cosf(th)));        /* :120-121, scalar expression */
          const float phM = ph + dA * mvY;
          float q[3], u[3];
          arr_sph_to_cart(r, thM, phM, pk, q);                     /* fromRotatedSphere :77-99 */
          mat3_tvec(R, q[0], q[1], q[2], u);
          erp_from_sphere_arr(c, u[0], u[1], u[2], &MX, &MY);
        }
      }
      else if (model == ORC_ROTATIONAL)
      {
        if (!mv_zero)
        {
          float s[3], u[3];
          erp_to_sphere_arr(c, X, Y, pk, s);
          mat3_vec(Rrot, s[0], s[1], s[2], u);                     /* :68-75 */
          erp_from_sphere_arr(c, u[0], u[1], u[2], &MX, &MY);
        }
      }
      else if (model == ORC_THREE_D_TRANSLATIONAL)
      {
        if (!mv_zero)
        {
          float s[3];
          erp_to_sphere_arr(c, X, Y, pk, s);
          erp_from_sphere_arr(c, s[0] + d3[0], s[1] + d3[1], s[2] + d3[2], &MX, &MY);
        }
      }
      else if (model == ORC_TANGENTIAL)
      {
        if (!mv_zero)
        {
          /* TangentialMotionModel.cpp:19-47 */
          float s[3], r, th, ph;
          erp_to_sphere_arr(c, X, Y, pk, s);
          arr_cart_to_sph(s[0], s[1], s[2], &r, &th, &ph);
          const float eps = PI_2_F - th;
          const float dAl = ph - alphaC;
          const float cosPsi = sinC * a_sin(eps, pk) + cosC * a_cos(eps, pk) * a_cos(dAl, pk);
          const float yPl = (a_sin(eps, pk) * cosC - sinC * a_cos(eps, pk) * a_cos(dAl, pk)) / cosPsi;
          const float xPl = (a_sin(dAl, pk) * a_cos(eps, pk)) / cosPsi;
          const float yM = yPl - mvY * dA;
          const float xM = xPl - mvX * dA;
          const float rho = sqrtf(xM * xM + yM * yM);
          const float eta = atanf(rho);
          const float gamma = rho * cosC * a_cos(eta, pk) - yM * sinC * a_sin(eta, pk);
          const float alphaM = alphaC + atanf((xM * sinf(eta)) / gamma);
          const float epsM = asinf(cosf(eta) * sinC + (yM * sinf(eta) * cosC) / rho);
          float q[3];
          arr_sph_to_cart(1.0f, PI_2_F - epsM, alphaM, pk, q);
          erp_from_sphere_arr(c, q[0], q[1], q[2], &MX, &MY);
        }
      }
      /* :152-165 NaN -> no motion; remove offset; chroma rescale; fixed point with std::round */
      const int nan = isnan(MX) || isnan(MY);
      float fx = (nan ? X : MX) - off, fy = (nan ? Y : MY) - off;
      if (!is_luma) { fx = fx / scaleX; fy = fy / scaleY; }
      out_x[i] = (int32_t)roundf(fx * (float)(1 << shiftHor));
      out_y[i] = (int32_t)roundf(fy * (float)(1 << shiftVer));
    }
}

/* ---------------------------------------------------------------------------------------------
 * InterpolationFilter (InterpolationFilter.cpp:82-100, 187-221, 392-645, 675-809)
 * ------------------------------------------------------------------------------------------- */
static const int16_t kLumaFilter[16][8] = {
  {0, 0, 0, 64, 0, 0, 0, 0},       {0, 1, -3, 63, 4, -2, 1, 0},     {-1, 2, -5, 62, 8, -3, 1, 0},
  {-1, 3, -8, 60, 13, -4, 1, 0},   {-1, 4, -10, 58, 17, -5, 1, 0},  {-1, 4, -11, 52, 26, -8, 3, -1},
  {-1, 3, -9, 47, 31, -10, 4, -1}, {-1, 4, -11, 45, 34, -10, 4, -1}, {-1, 4, -11, 40, 40, -11, 4, -1},
  {-1, 4, -10, 34, 45, -11, 4, -1}, {-1, 4, -10, 31, 47, -9, 3, -1}, {-1, 3, -8, 26, 52, -11, 4, -1},
  {0, 1, -5, 17, 58, -10, 4, -1},  {0, 1, -4, 13, 60, -8, 3, -1},   {0, 1, -3, 8, 62, -5, 2, -1},
  {0, 1, -2, 4, 63, -3, 1, 0}};
static const int16_t kChromaFilter[32][4] = {
  {0, 64, 0, 0},   {-1, 63, 2, 0},  {-2, 62, 4, 0},  {-2, 60, 7, -1},  {-2, 58, 10, -2}, {-3, 57, 12, -2},
  {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
  {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
  {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
  {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
  {0, 4, 62, -2},  {0, 2, 63, -1}};

#define IF_INTERNAL_PREC 14
#define IF_FILTER_PREC 6
#define IF_INTERNAL_OFFS (1 << (IF_INTERNAL_PREC - 1))
static int if_frac_bits(int bd) { int v = IF_INTERNAL_PREC - bd; return v < 2 ? 2 : v; } /* max(2, 14-bd) */
static int16_t clip_pel(int v, int lo, int hi) { return (int16_t)(v < lo ? lo : (v > hi ? hi : v)); }

/* filterCopy<isFirst,isLast> (:392-538), no DMVR-bilinear */
static void filter_copy(const int16_t* src, int ss, int16_t* dst, int ds, int w, int h, int is_first, int is_last,
                        int bd, int lo, int hi)
{
  const int shift = if_frac_bits(bd);
  for (int r = 0; r < h; ++r, src += ss, dst += ds)
    for (int x = 0; x < w; ++x)
    {
      if (is_first == is_last) dst[x] = src[x];
      else if (is_first)
      {
        int16_t val = (int16_t)(src[x] << shift); /* leftShift_round */
        dst[x] = (int16_t)(val - (int16_t)IF_INTERNAL_OFFS);
      }
      else
      {
        int16_t val = src[x];
        val = (int16_t)((val + IF_INTERNAL_OFFS + (1 << (shift - 1))) >> shift); /* rightShift_round */
        dst[x] = clip_pel(val, lo, hi);
      }
    }
}
/* filter<N,isVertical,isFirst,isLast> (:540-645) */
static void filter_n(int N, int vertical, int is_first, int is_last, const int16_t* src, int ss, int16_t* dst,
                     int ds, int w, int h, const int16_t* coeff, int bd, int lo, int hi)
{
  const int cs = vertical ? ss : 1;
  src -= (N / 2 - 1) * cs;
  const int headRoom = if_frac_bits(bd);
  int shift = IF_FILTER_PREC, offset;
  if (is_last)
  {
    shift += is_first ? 0 : headRoom;
    offset = 1 << (shift - 1);
    offset += is_first ? 0 : IF_INTERNAL_OFFS << IF_FILTER_PREC;
  }
  else
  {
    shift -= is_first ? headRoom : 0;
    offset = is_first ? -IF_INTERNAL_OFFS * (1 << shift) : 0;
  }
  for (int r = 0; r < h; ++r, src += ss, dst += ds)
    for (int x = 0; x < w; ++x)
    {
      int sum = 0;
      for (int k = 0; k < N; ++k) sum += src[x + k * cs] * coeff[k];
      int16_t val = (int16_t)((sum + offset) >> shift);
      if (is_last) val = clip_pel(val, lo, hi);
      dst[x] = val;
    }
}
void orc_filter_hor(int is_luma, const int16_t* src, int ss, int16_t* dst, int ds, int w, int h, int frac,
                    int is_last, int bd, int lo, int hi)
{
  if (frac == 0) filter_copy(src, ss, dst, ds, w, h, 1, is_last, bd, lo, hi);
  else if (is_luma) filter_n(8, 0, 1, is_last, src, ss, dst, ds, w, h, kLumaFilter[frac], bd, lo, hi);
  else filter_n(4, 0, 1, is_last, src, ss, dst, ds, w, h, kChromaFilter[frac], bd, lo, hi);
}
void orc_filter_ver(int is_luma, const int16_t* src, int ss, int16_t* dst, int ds, int w, int h, int frac,
                    int is_first, int is_last, int bd, int lo, int hi)
{
  if (frac == 0) filter_copy(src, ss, dst, ds, w, h, is_first, is_last, bd, lo, hi);
  else if (is_luma) filter_n(8, 1, is_first, is_last, src, ss, dst, ds, w, h, kLumaFilter[frac], bd, lo, hi);
  else filter_n(4, 1, is_first, is_last, src, ss, dst, ds, w, h, kChromaFilter[frac], bd, lo, hi);
}

/* one sub-block, shared by xPredInterBlkMM (:773-828) and xMVReprojectionInterpolation (:6330-6375) */
static void mc_subblock(const int16_t* ref, int rs, int is_luma, int xPos, int yPos, int xFrac, int yFrac, int sbw,
                        int sbh, int rnd, int bd, int lo, int hi, int16_t* dst, int ds)
{
  const int16_t* p = ref + (long)yPos * rs + xPos;
  if (yFrac == 0) orc_filter_hor(is_luma, p, rs, dst, ds, sbw, sbh, xFrac, rnd, bd, lo, hi);
  else if (xFrac == 0) orc_filter_ver(is_luma, p, rs, dst, ds, sbw, sbh, yFrac, 1, rnd, bd, lo, hi);
  else
  {
    int16_t tmp[16 * 16];
    const int taps = is_luma ? 8 : 4;
    orc_filter_hor(is_luma, p - ((taps >> 1) - 1) * rs, rs, tmp, 16, sbw, sbh + taps - 1, xFrac, 0, bd, lo, hi);
    orc_filter_ver(is_luma, tmp + ((taps >> 1) - 1) * 16, 16, dst, ds, sbw, sbh, yFrac, 0, rnd, bd, lo, hi);
  }
}

void orc_mc_block(OrcCtx* c, const int16_t* ref, int rs, int comp, int chroma_fmt, int pos_x, int pos_y, int size_w,
                  int size_h, int mv_hor, int mv_ver, int model, int bi, int lo, int hi, int oob_mode, int16_t* dst,
                  int ds)
{
  /* the ME twin passes CHROMA_400 as a dummy format and only does luma (InterSearch.cpp:6299) */
  const int csx = scale_x(comp, chroma_fmt), csy = scale_y(comp, chroma_fmt);
  const int sbw = 4 >> csx, sbh = 4 >> csy, shH = 4 + csx, shV = 4 + csy;
  const int rows = size_h / sbh, cols = size_w / sbw;
  int32_t* px = (int32_t*)malloc(sizeof(int32_t) * rows * cols * 2);
  int32_t* py = px + rows * cols;
  orc_reproject(c, pos_x, pos_y, size_w, size_h, mv_hor, mv_ver, model, comp, chroma_fmt, px, py);
  const int picw = c->cfg.pic_w >> csx, pich = c->cfg.pic_h >> csy; /* refBuf.width/height of the component */
  const int mcw = (oob_mode == ORC_OOB_ME) ? 0 : c->cfg.max_cu / (1 << csx);
  const int mch = (oob_mode == ORC_OOB_ME) ? 0 : c->cfg.max_cu / (1 << csy);
  const int rnd = !bi;
  for (int col = 0; col < cols; ++col)
    for (int row = 0; row < rows; ++row)
    {
      const int i = col * rows + row;
      const int xP = px[i] >> shH, yP = py[i] >> shV, xF = px[i] & ((1 << shH) - 1), yF = py[i] & ((1 << shV) - 1);
      int16_t* d = dst + row * sbh * ds + col * sbw;
      /* InterPrediction.cpp:780 / InterSearch.cpp:6332.  refBuf.width/height and subblockSize.width/height are SizeType
         (uint32_t), so the two upper-bound comparisons are UNSIGNED: a negative position converts to a huge value and
         counts as beyond the bound.  Net effect: the +-maxCU slack exists on the right/bottom only; any sub-block whose
         integer position is negative is zero-filled.  (Found by decoding a real bitstream: tests/test_e2e_*.) */
      if (xP < -mcw || yP < -mch || (uint32_t)xP >= (uint32_t)(picw + mcw - sbw) || (uint32_t)yP >= (uint32_t)(pich + mch - sbh))
      {
        for (int y = 0; y < sbh; ++y) memset(d + y * ds, 0, sbw * sizeof(int16_t));
        continue;
      }
      mc_subblock(ref, rs, comp == 0, xP, yP, xF, yF, sbw, sbh, rnd, c->cfg.bit_depth, lo, hi, d, ds);
    }
  free(px);
}

void orc_mc_block_classic(const int16_t* ref, int rs, int comp, int chroma_fmt, int pos_x, int pos_y, int size_w,
                          int size_h, int mv_hor, int mv_ver, int bi, int bd, int lo, int hi, int16_t* dst, int ds)
{
  /* InterPrediction.cpp:928-1096 without wrap/RPR/BDOF/DMVR: shift = MV_FRACTIONAL_BITS_INTERNAL + scale,
   * xFrac = mv.hor & mask, yFrac likewise, refBuf offset = mv >> shift; filterHor/filterVer as for MM. */
  const int csx = scale_x(comp, chroma_fmt), csy = scale_y(comp, chroma_fmt);
  const int shH = 4 + csx, shV = 4 + csy;
  /* :960-971: luma phase = mv & 15; chroma phase = mv * (1 << (1 - scale)) & 31 -- the chroma table always has 32 phases,
   * so for an un-subsampled direction (4:4:4, 4:2:2 vertical) the 1/16 fraction is doubled */
  const int xF = (comp == 0) ? (mv_hor & 15) : ((mv_hor * (1 << (1 - csx))) & 31);
  const int yF = (comp == 0) ? (mv_ver & 15) : ((mv_ver * (1 << (1 - csy))) & 31);
  const int xI = mv_hor >> shH, yI = mv_ver >> shV;
  const int is_luma = comp == 0, rnd = !bi, taps = is_luma ? 8 : 4;
  const int16_t* p = ref + (long)(pos_y + yI) * rs + (pos_x + xI);
  if (yF == 0) orc_filter_hor(is_luma, p, rs, dst, ds, size_w, size_h, xF, rnd, bd, lo, hi);
  else if (xF == 0) orc_filter_ver(is_luma, p, rs, dst, ds, size_w, size_h, yF, 1, rnd, bd, lo, hi);
  else
  {
    int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * size_w * (size_h + taps - 1));
    orc_filter_hor(is_luma, p - ((taps >> 1) - 1) * rs, rs, tmp, size_w, size_w, size_h + taps - 1, xF, 0, bd, lo, hi);
    orc_filter_ver(is_luma, tmp + ((taps >> 1) - 1) * size_w, size_w, dst, ds, size_w, size_h, yF, 0, rnd, bd, lo, hi);
    free(tmp);
  }
}

void orc_mc_bdof(const int16_t* ref, int rs, int pos_x, int pos_y, int size_w, int size_h, int mv_hor, int mv_ver, int bd, int lo,
                 int hi, int16_t* tmp)
{
  /* xPredInterBlk with bioApplied, luma (InterPrediction.cpp:1007-1015, 1054-1093): the 14-bit prediction goes to (2,2) of a
   * buffer of stride width + 2*BIO_EXTEND_SIZE + 2 = w + 4; the ring around it holds integer-position reference samples
   * (leftShift_round(ref, IF_INTERNAL_FRAC_BITS) - IF_INTERNAL_OFFS), fetched one sample up/left when the fraction is < 8. */
  const int stride = size_w + 4, width = size_w + 4, height = size_h + 4;
  orc_mc_block_classic(ref, rs, 0, 0, pos_x, pos_y, size_w, size_h, mv_hor, mv_ver, 1, bd, lo, hi, tmp + 2 * stride + 2, stride);
  const int shift = 14 - bd < 2 ? 2 : 14 - bd;
  const int xFrac = mv_hor & 15, yFrac = mv_ver & 15;
  const int xOffset = (xFrac < 8) ? 1 : 0, yOffset = (yFrac < 8) ? 1 : 0;
  const int16_t* refBuf = ref + (long)(pos_y + (mv_ver >> 4)) * rs + (pos_x + (mv_hor >> 4));
  const int16_t* refPel = refBuf - yOffset * rs - xOffset;
  int16_t* dstPel = tmp + stride + 1;
  for (int w = 0; w < width - 2; w++) dstPel[w] = (int16_t)((int16_t)(refPel[w] << shift) - (int16_t)8192);
  refPel = refBuf + (1 - yOffset) * rs - xOffset;
  dstPel = tmp + 2 * stride + 1;
  for (int h = 0; h < height - 2 - 2; h++)
  {
    dstPel[0] = (int16_t)((int16_t)(refPel[0] << shift) - (int16_t)8192);
    dstPel[width - 3] = (int16_t)((int16_t)(refPel[width - 3] << shift) - (int16_t)8192);
    refPel += rs;
    dstPel += stride;
  }
  refPel = refBuf + (height - 2 - 2 + 1 - yOffset) * rs - xOffset;
  dstPel = tmp + (height - 2) * stride + 1;
  for (int w = 0; w < width - 2; w++) dstPel[w] = (int16_t)((int16_t)(refPel[w] << shift) - (int16_t)8192);
}

/* ---------------------------------------------------------------------------------------------
 * RdCost SAD / HAD
 * ------------------------------------------------------------------------------------------- */
uint64_t orc_sad(const int16_t* org, int os, const int16_t* cur, int cs, int w, int h, int sub_shift)
{
  /* RdCost.cpp:482-520 with maximumDistortionForEarlyExit = max (the sized SIMD SADs ignore it,
   * x86/RdCostX86.h:378-495); DISTORTION_PRECISION_ADJUSTMENT = 0 (FULL_NBIT, TypeDef.h:235-249) */
  const int step = 1 << sub_shift;
  uint64_t sum = 0;
  for (int y = 0; y < h; y += step)
    for (int x = 0; x < w; ++x) sum += (uint64_t)abs(org[y * os + x] - cur[y * cs + x]);
  return sum << sub_shift;
}

/* Generic separable Hadamard of a tw x th tile of differences.  Sum of |coefficients| does not
 * depend on butterfly order or coefficient sign; only the DC term is treated specially
 * (JVET_R0164_MEAN_SCALED_SATD, TypeDef.h:74), so the explicit butterflies of RdCost.cpp:2264-2941
 * collapse to this.  The identity is checked in tests against a literal transcription of 4x4. */
static void hadamard_1d(int* v, int n, int stride)
{
  for (int len = 1; len < n; len <<= 1)
    for (int i = 0; i < n; i += len << 1)
      for (int j = i; j < i + len; ++j)
      {
        int a = v[j * stride], b = v[(j + len) * stride];
        v[j * stride] = a + b;
        v[(j + len) * stride] = a - b;
      }
}
static uint64_t had_tile(const int16_t* org, int os, const int16_t* cur, int cs, int tw, int th, int had_mode)
{
  int d[16 * 16];
  for (int y = 0; y < th; ++y)
    for (int x = 0; x < tw; ++x) d[y * tw + x] = org[y * os + x] - cur[y * cs + x];
  for (int y = 0; y < th; ++y) hadamard_1d(d + y * tw, tw, 1);
  for (int x = 0; x < tw; ++x) hadamard_1d(d + x, th, tw);
  uint64_t sad = 0;
  for (int i = 0; i < tw * th; ++i) sad += (uint64_t)abs(d[i]);
  const uint64_t dc = (uint64_t)abs(d[0]);
  sad -= dc;
  sad += dc >> 2;
  if (tw == 2 && th == 2) return sad;                     /* xCalcHADs2x2 :2264-2288 */
  if (tw == 4 && th == 4) return (sad + 1) >> 1;          /* :2386 */
  if (tw == 8 && th == 8) return (sad + 2) >> 2;          /* :2487 */
  if (had_mode == ORC_HAD_SIMD)
  {
    /* x86/RdCostX86.h:2209, 2661-2662 (16x8), 2860-2861 (8x16), 3011-3012 (4x8), 3160-3161 (8x4) */
    uint32_t s32 = (uint32_t)sad;
    s32 = (uint32_t)(((uint64_t)s32 * 0xb504f334ULL) >> 32);
    return (tw * th == 128) ? (s32 >> 2) : (s32 >> 1);
  }
  if (tw * th == 128) return (uint64_t)(int)((double)sad / sqrt(16.0 * 8) * 2); /* :2637, :2778 */
  return (uint64_t)(int)((double)sad / sqrt(4.0 * 8) * 2);                      /* :2855, :2938 */
}
uint64_t orc_had(const int16_t* org, int os, const int16_t* cur, int cs, int w, int h, int had_mode)
{
  /* tiling rule: RdCost.cpp:2962-3060 == x86/RdCostX86.h:4434-4553 for power-of-two sizes
   * (the AVX2 16x16 case is four 8x8 results added, :3165-3340). */
  int tw, th;
  if (w > h && (h & 7) == 0 && (w & 15) == 0) { tw = 16; th = 8; }
  else if (w < h && (w & 7) == 0 && (h & 15) == 0) { tw = 8; th = 16; }
  else if (w > h && (h & 3) == 0 && (w & 7) == 0) { tw = 8; th = 4; }
  else if (w < h && (w & 3) == 0 && (h & 7) == 0) { tw = 4; th = 8; }
  else if ((h % 8 == 0) && (w % 8 == 0)) { tw = 8; th = 8; }
  else if ((h % 4 == 0) && (w % 4 == 0)) { tw = 4; th = 4; }
  else { tw = 2; th = 2; }
  uint64_t sum = 0;
  for (int y = 0; y < h; y += th)
    for (int x = 0; x < w; x += tw) sum += had_tile(org + y * os + x, os, cur + y * cs + x, cs, tw, th, had_mode);
  return sum;
}

void orc_me_block(OrcCtx* c, const int16_t* ref, int rs, const int16_t* org, int os, int pos_x, int pos_y, int size_w,
                  int size_h, int model, int n, const int32_t* mv_hor, const int32_t* mv_ver, int mv_prec_shift,
                  int dist_kind, int sub_shift, int had_mode, int lo, int hi, uint64_t* dist)
{
  int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * 128 * 128); /* m_tmpMMStorage, InterSearch.cpp:292 */
  for (int i = 0; i < n; ++i)
  {
    /* Mv::changePrecision(mvPrec, MV_PRECISION_INTERNAL): left shift */
    const int mh = mv_hor[i] * (1 << mv_prec_shift), mv = mv_ver[i] * (1 << mv_prec_shift);
    if (model == ORC_CLASSIC)
      orc_mc_block_classic(ref, rs, 0, ORC_CHROMA_400, pos_x, pos_y, size_w, size_h, mh, mv, 0, c->cfg.bit_depth, lo, hi, tmp, 128);
    else
      orc_mc_block(c, ref, rs, 0, ORC_CHROMA_400, pos_x, pos_y, size_w, size_h, mh, mv, model, 0, lo, hi, ORC_OOB_ME, tmp, 128);
    dist[i] = (dist_kind == ORC_DIST_SAD) ? orc_sad(org, os, tmp, 128, size_w, size_h, sub_shift)
                                          : orc_had(org, os, tmp, 128, size_w, size_h, had_mode);
  }
  free(tmp);
}

void orc_add_avg(const int16_t* s0, int st0, const int16_t* s1, int st1, int16_t* dst, int ds, int w, int h, int bd,
                 int lo, int hi)
{
  /* Buffer.cpp addAvg: shift = max(2, 14-bd)+1, offset = (1<<(shift-1)) + 2*IF_INTERNAL_OFFS */
  const int shift = if_frac_bits(bd) + 1;
  const int offset = (1 << (shift - 1)) + 2 * IF_INTERNAL_OFFS;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      dst[y * ds + x] = clip_pel((s0[y * st0 + x] + s1[y * st1 + x] + offset) >> shift, lo, hi);
}

void orc_extend_border(int16_t* p, int stride, int w, int h, int mx, int my)
{
  /* Picture.cpp:1008-1046: left/right replicate per row, then copy first/last row into the margins */
  for (int y = 0; y < h; ++y)
  {
    int16_t* r = p + (long)y * stride;
    for (int x = 0; x < mx; ++x) { r[-mx + x] = r[0]; r[w + x] = r[w - 1]; }
  }
  int16_t* last = p + (long)(h - 1) * stride - mx;
  for (int y = 0; y < my; ++y) memcpy(last + (long)(y + 1) * stride, last, sizeof(int16_t) * (w + 2 * mx));
  int16_t* first = p - mx;
  for (int y = 0; y < my; ++y) memcpy(first - (long)(y + 1) * stride, first, sizeof(int16_t) * (w + 2 * mx));
}

void orc_extend_wrap_border(int16_t* p, int stride, int w, int h, int mx, int my, int xoffset)
{
  /* Picture::extendWrapBorder, Picture.cpp:1049-1092 (loop order as in the reference) */
  int16_t* pi = p;
  for (int y = 0; y < h; ++y)
  {
    for (int x = 0; x < mx; ++x)
    {
      if (x < xoffset)
      {
        pi[-x - 1] = pi[-x - 1 + xoffset];
        pi[w + x] = pi[w + x - xoffset];
      }
      else
      {
        pi[-x - 1] = pi[0];
        pi[w + x] = pi[w - 1];
      }
    }
    pi += stride;
  }
  pi -= (stride + mx);
  for (int y = 0; y < my; ++y) memcpy(pi + (long)(y + 1) * stride, pi, sizeof(int16_t) * (w + (mx << 1)));
  pi -= (long)(h - 1) * stride;
  for (int y = 0; y < my; ++y) memcpy(pi - (long)(y + 1) * stride, pi, sizeof(int16_t) * (w + (mx << 1)));
}

/* ---------------------------------------------------------------------------------------------
 * DMVR (projected): InterPrediction.cpp:2442-2635, 2147-2155, 1958-2048, 2157-2175
 * ------------------------------------------------------------------------------------------- */
static const int kSearchOffset[25][2] = { /* InterPrediction.h:107-111 (hor, ver) */
  {-2, -2}, {-1, -2}, {0, -2}, {1, -2}, {2, -2}, {-2, -1}, {-1, -1}, {0, -1}, {1, -1}, {2, -1},
  {-2, 0},  {-1, 0},  {0, 0},  {1, 0},  {2, 0},  {-2, 1},  {-1, 1},  {0, 1},  {1, 1},  {2, 1},
  {-2, 2},  {-1, 2},  {0, 2},  {1, 2},  {2, 2}};

static uint64_t dmvr_cost(const int16_t* a, int as, const int16_t* b, int bs, int w, int h, int simd)
{
  /* xDMVRCost: setDistParam(..., subShiftMode 1) -> DF_SAD + log2(width); return distFunc >> 1 */
  uint64_t sum = 0;
  for (int y = 0; y < h; y += 2)
    for (int x = 0; x < w; ++x)
    {
      int d = a[y * as + x] - b[y * bs + x];
      if (simd) { int16_t d16 = (int16_t)d; int ad = d16 < 0 ? -(int)d16 : (int)d16; sum += (uint64_t)(ad & 0xffff); }
      else sum += (uint64_t)abs(d);
    }
  sum <<= 1;
  return sum >> 1;
}

void orc_dmvr_costs(OrcCtx* c, const int16_t* ref0, int rs0, const int16_t* ref1, int rs1, int x, int y, int dx, int dy,
                    int mv0h, int mv0v, int mv1h, int mv1v, int model, const float* R0, const float* R1, int lo, int hi,
                    int simd, uint64_t costs25[25])
{
  int16_t p0[16 * 16], p1[16 * 16];
  for (int k = 0; k < 25; ++k)
  {
    const int oh = kSearchOffset[k][0] * 16, ov = kSearchOffset[k][1] * 16; /* << MV_FRACTIONAL_BITS_INTERNAL */
    if (R0) orc_set_rotation(c, model, R0);
    orc_mc_block(c, ref0, rs0, 0, c->cfg.chroma_fmt, x, y, dx, dy, mv0h + oh, mv0v + ov, model, 1, lo, hi, ORC_OOB_PRED, p0, 16);
    if (R1) orc_set_rotation(c, model, R1);
    orc_mc_block(c, ref1, rs1, 0, c->cfg.chroma_fmt, x, y, dx, dy, mv1h - oh, mv1v - ov, model, 1, lo, hi, ORC_OOB_PRED, p1, 16);
    costs25[k] = dmvr_cost(p0, 16, p1, 16, dx, dy, simd);
  }
}

/* ---------------------------------------------------------------------------------------------
 * DMVR (classic): InterPrediction::xProcessDMVR (InterPrediction.cpp:2239-2439) -- literal: a prefetched window in a
 * small buffer (xPrefetch :1891-1936), padded by replication (xPad :1938-1956), bilinear search predictions from it
 * (xinitMC :2177-2223, FILTER_DMVR = m_bilinearFilterPrec4, 10-bit output), 8-tap final MC from it
 * (xFinalPaddedMCForDMVR :2075-2145).
 * ------------------------------------------------------------------------------------------- */
static const int16_t kLumaAltHpel[8] = {0, 3, 9, 20, 20, 9, 3, 0};  /* InterpolationFilter.cpp:186 */

/* filterCopy<true,false> / filter<2,*,isFirst,false> with biMCForDMVR (InterpolationFilter.cpp:416-447, 588-600) */
static void filter_bilinear(int vertical, int is_first, const int16_t* src, int ss, int16_t* dst, int ds, int w, int h, int frac, int bd)
{
  if (frac == 0)  /* only reached as filterHor's copy: isFirst = true */
  {
    for (int r = 0; r < h; ++r, src += ss, dst += ds)
      for (int x = 0; x < w; ++x)
      {
        if (bd - 10 > 0) dst[x] = (int16_t)((src[x] + (1 << (bd - 10 - 1))) >> (bd - 10));
        else dst[x] = (int16_t)(src[x] << (10 - bd));
      }
    return;
  }
  const int cs = vertical ? ss : 1;
  const int c0 = 16 - frac, c1 = frac;                       /* m_bilinearFilterPrec4[frac] (:297-315) */
  const int shift = is_first ? 4 - (10 - bd) : 4, offset = 1 << (shift - 1);
  for (int r = 0; r < h; ++r, src += ss, dst += ds)
    for (int x = 0; x < w; ++x) dst[x] = (int16_t)((src[x] * c0 + src[x + cs] * c1 + offset) >> shift);
}

/* window of the reference plane -> pad buffer at (2,2) (xPrefetch: copyBuffer of (w + taps-1) x (h + taps-1)) */
static void dmvr_prefetch(const int16_t* ref, int rs, int ox, int oy, int ww, int wh, int16_t* pad, int S)
{
  for (int y = 0; y < wh; ++y)
    for (int x = 0; x < ww; ++x) pad[(2 + y) * S + 2 + x] = ref[(long)(oy + y) * rs + ox + x];
}
/* g_pelBufOP.padding (Buffer.cpp paddingCore): left/right columns, then whole rows up and down */
static void dmvr_pad(int16_t* pad, int S, int ww, int wh, int padsize)
{
  int16_t* p = pad + 2 * S + 2;
  for (int y = 0; y < wh; ++y)
    for (int k = 1; k <= padsize; ++k) { p[y * S - k] = p[y * S]; p[y * S + ww - 1 + k] = p[y * S + ww - 1]; }
  for (int k = 1; k <= padsize; ++k)
    for (int x = -padsize; x < ww + padsize; ++x) { p[-k * S + x] = p[x]; p[(wh - 1 + k) * S + x] = p[(wh - 1) * S + x]; }
}

void orc_dmvr_classic_costs(const int16_t* ref0, int rs0, const int16_t* ref1, int rs1, int ox0, int oy0, int fx0, int fy0,
                            int ox1, int oy1, int fx1, int fy1, int dx, int dy, int bd, uint64_t costs25[25])
{
  /* (ox, oy): xPrefetch's Rec_offset of the list -- luma position of the first prefetched sample, i.e.
   * pu.lumaPos() + (clipMv(mergeMv - (3 << 4)) >> 4); (fx, fy) = clipMv(mergeMv) & 15 (xinitMC, xPredInterBlk :965-966). */
  const int S = dx + 4 + 8, PS = dx + 4;   /* pcPad stride (:1907), m_biLinearBufStride (:2307) */
  int16_t pad[2][(16 + 12) * (16 + 12)], pred[2][20 * 20], tmp[20 * 21];
  const int16_t* ref[2] = {ref0, ref1};
  const int rs[2] = {rs0, rs1}, ox[2] = {ox0, ox1}, oy[2] = {oy0, oy1}, fx[2] = {fx0, fx1}, fy[2] = {fy0, fy1};
  for (int l = 0; l < 2; ++l)
  {
    memset(pad[l], 0, sizeof(pad[l]));
    dmvr_prefetch(ref[l], rs[l], ox[l], oy[l], dx + 7, dy + 7, pad[l], S);
    /* xinitMC: offset = (2 + 3) * (S + 1) - 2 * S - 2; xPredInterBlk(bi = true, dmvrWidth = dx + 4, bilinearMC) (:1019-1049) */
    const int16_t* src = pad[l] + 3 * (S + 1);
    const int w = dx + 4, h = dy + 4;
    if (fy[l] == 0) filter_bilinear(0, 1, src, S, pred[l], PS, w, h, fx[l], bd);
    else if (fx[l] == 0) filter_bilinear(1, 1, src, S, pred[l], PS, w, h, fy[l], bd);
    else
    {
      filter_bilinear(0, 1, src, S, tmp, w, w, h + 1, fx[l], bd);   /* tmpBuf = PelBuf(.., Size(dmvrWidth, dmvrHeight)): stride w */
      filter_bilinear(1, 0, tmp, w, pred[l], PS, w, h, fy[l], bd);
    }
  }
  for (int k = 0; k < 25; ++k)
  {
    const int oh = kSearchOffset[k][0], ov = kSearchOffset[k][1];  /* xBIPMVRefine (:2050-2073): L0 + offset, L1 - offset */
    costs25[k] = dmvr_cost(pred[0] + (2 + ov) * PS + 2 + oh, PS, pred[1] + (2 - ov) * PS + 2 - oh, PS, dx, dy, 0);
  }
}

/* the 14-bit prediction + integer-sample ring of xPredInterBlk with bioApplied (:1007-1015, 1054-1093) from any source pointer */
static void bdof_ring(const int16_t* refBuf, int rs, int size_w, int size_h, int xFrac, int yFrac, int bd, int16_t* tmp)
{
  const int stride = size_w + 4, width = size_w + 4, height = size_h + 4;
  const int shift = 14 - bd < 2 ? 2 : 14 - bd;
  const int xOffset = (xFrac < 8) ? 1 : 0, yOffset = (yFrac < 8) ? 1 : 0;
  const int16_t* refPel = refBuf - yOffset * rs - xOffset;
  int16_t* dstPel = tmp + stride + 1;
  for (int w = 0; w < width - 2; w++) dstPel[w] = (int16_t)((int16_t)(refPel[w] << shift) - (int16_t)8192);
  refPel = refBuf + (1 - yOffset) * rs - xOffset;
  dstPel = tmp + 2 * stride + 1;
  for (int h = 0; h < height - 2 - 2; h++)
  {
    dstPel[0] = (int16_t)((int16_t)(refPel[0] << shift) - (int16_t)8192);
    dstPel[width - 3] = (int16_t)((int16_t)(refPel[width - 3] << shift) - (int16_t)8192);
    refPel += rs;
    dstPel += stride;
  }
  refPel = refBuf + (height - 2 - 2 + 1 - yOffset) * rs - xOffset;
  dstPel = tmp + (height - 2) * stride + 1;
  for (int w = 0; w < width - 2; w++) dstPel[w] = (int16_t)((int16_t)(refPel[w] << shift) - (int16_t)8192);
}

/* xPredInterBlk's filter cascade (:1019-1049) from any source pointer, default filters, optional alternative half-pel filter */
static void mc_from_ptr(const int16_t* p, int rs, int is_luma, int xF, int yF, int w, int h, int rnd, int alt, int bd, int lo, int hi,
                        int16_t* dst, int ds)
{
  const int taps = is_luma ? 8 : 4;
  const int altx = alt && is_luma && xF == 8, alty = alt && is_luma && yF == 8;   /* InterpolationFilter.cpp:713-716, 781-784 */
  if (yF == 0)
  {
    if (altx) filter_n(8, 0, 1, rnd, p, rs, dst, ds, w, h, kLumaAltHpel, bd, lo, hi);
    else orc_filter_hor(is_luma, p, rs, dst, ds, w, h, xF, rnd, bd, lo, hi);
  }
  else if (xF == 0)
  {
    if (alty) filter_n(8, 1, 1, rnd, p, rs, dst, ds, w, h, kLumaAltHpel, bd, lo, hi);
    else orc_filter_ver(is_luma, p, rs, dst, ds, w, h, yF, 1, rnd, bd, lo, hi);
  }
  else
  {
    int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * w * (h + taps - 1));
    const int16_t* q = p - ((taps >> 1) - 1) * rs;
    if (altx) filter_n(8, 0, 1, 0, q, rs, tmp, w, w, h + taps - 1, kLumaAltHpel, bd, lo, hi);
    else orc_filter_hor(is_luma, q, rs, tmp, w, w, h + taps - 1, xF, 0, bd, lo, hi);
    if (alty) filter_n(8, 1, 0, rnd, tmp + ((taps >> 1) - 1) * w, w, dst, ds, w, h, kLumaAltHpel, bd, lo, hi);
    else orc_filter_ver(is_luma, tmp + ((taps >> 1) - 1) * w, w, dst, ds, w, h, yF, 0, rnd, bd, lo, hi);
    free(tmp);
  }
}

void orc_dmvr_final_mc(const int16_t* ref, int rs, int is_luma, int ox, int oy, int w, int h, int padsize, int dix, int diy,
                       int xFrac, int yFrac, int alt_hpel, int bio, int bd, int lo, int hi, int16_t* dst, int ds)
{
  /* one component of one list of xFinalPaddedMCForDMVR for a sub-PU that moved (or luma of one that did not):
   * (ox, oy) = xPrefetch's Rec_offset in component samples, (w, h) the component block, padsize = 2 >> scaleY (:1950),
   * (dix, diy) = deltaIntMv (:2126-2127), fractions from cMvClipped; output as xPredInterBlk(bi = true): 14-bit, and with
   * bio (luma) into (2,2) of the (w + 4)-stride BDOF buffer with its ring. */
  const int taps = is_luma ? 8 : 4, left = (taps >> 1) - 1;
  const int S = w + 4 + 8;
  int16_t* pad = (int16_t*)calloc((size_t)S * (h + 4 + 8), sizeof(int16_t));
  dmvr_prefetch(ref, rs, ox, oy, w + taps - 1, h + taps - 1, pad, S);
  dmvr_pad(pad, S, w + taps - 1, h + taps - 1, padsize);
  const int16_t* src = pad + (2 + left) * (S + 1) + diy * S + dix;   /* :2131-2134 */
  if (bio && is_luma)
  {
    const int stride = w + 4;
    mc_from_ptr(src, S, 1, xFrac, yFrac, w, h, 0, alt_hpel, bd, lo, hi, dst + 2 * stride + 2, stride);
    bdof_ring(src, S, w, h, xFrac, yFrac, bd, dst);
  }
  else mc_from_ptr(src, S, is_luma, xFrac, yFrac, w, h, 0, alt_hpel, bd, lo, hi, dst, ds);
  free(pad);
}

/* xPredInterBlk from the reference picture with pu.cu->imv == IMV_HPEL (:948): the alternative half-pel luma filter */
void orc_mc_block_classic_alt(const int16_t* ref, int rs, int pos_x, int pos_y, int size_w, int size_h, int mv_hor, int mv_ver, int bi,
                              int bd, int lo, int hi, int16_t* dst, int ds)
{
  const int16_t* p = ref + (long)(pos_y + (mv_ver >> 4)) * rs + (pos_x + (mv_hor >> 4));
  mc_from_ptr(p, rs, 1, mv_hor & 15, mv_ver & 15, size_w, size_h, !bi, 1, bd, lo, hi, dst, ds);
}

static int32_t div_for_maxq7(int64_t N, int64_t D)
{
  int32_t sign = 0, q = 0;
  if (N < 0) { sign = 1; N = -N; }
  D = (D << 3);
  if (N >= D) { N -= D; q++; }
  q = (q << 1);
  D = (D >> 1);
  if (N >= D) { N -= D; q++; }
  q = (q << 1);
  if (N >= (D >> 1)) q++;
  return sign ? -q : q;
}
static void sub_pel_error_srfc(const uint64_t* sb, int32_t* dmv)
{
  int64_t num, den;
  num = (int64_t)((sb[1] - sb[3]) << 4);
  den = (int64_t)((sb[1] + sb[3] - (sb[0] << 1)));
  if (0 != den)
  {
    if ((sb[1] != sb[0]) && (sb[3] != sb[0])) dmv[0] = div_for_maxq7(num, den);
    else dmv[0] = (sb[1] == sb[0]) ? -8 : 8;
  }
  num = (int64_t)((sb[2] - sb[4]) << 4);
  den = (int64_t)((sb[2] + sb[4] - (sb[0] << 1)));
  if (0 != den)
  {
    if ((sb[2] != sb[0]) && (sb[4] != sb[0])) dmv[1] = div_for_maxq7(num, den);
    else dmv[1] = (sb[2] == sb[0]) ? -8 : 8;
  }
}

int orc_dmvr_decide(const uint64_t costs25[25], int dx, int dy, int32_t delta_mv[2], uint64_t* min_cost)
{
  uint64_t sads[25];
  for (int i = 0; i < 25; ++i) sads[i] = UINT64_MAX;
  uint64_t* p = &sads[12];
  int notZero = 1;
  int16_t total[2] = {0, 0}, delta[2] = {0, 0};
  uint64_t minCost = costs25[12];
  minCost -= (minCost >> 2);
  if (minCost < (uint64_t)(dx * dy)) notZero = 0;
  else
  {
    p[0] = minCost;
    if (!minCost) notZero = 0;
    else
    {
      for (int n = 0; n < 25; ++n)
      {
        const int off = kSearchOffset[n][1] * 5 + kSearchOffset[n][0];
        if (p[off] == UINT64_MAX) p[off] = costs25[n];
        if (p[off] < minCost) { minCost = p[off]; delta[0] = (int16_t)kSearchOffset[n][0]; delta[1] = (int16_t)kSearchOffset[n][1]; }
      }
      if (!(delta[0] == 0 && delta[1] == 0))
      {
        total[0] += delta[0]; total[1] += delta[1];
        p += delta[1] * 5 + delta[0];
      }
    }
  }
  total[0] = (int16_t)(total[0] << 4);
  total[1] = (int16_t)(total[1] << 4);
  if (notZero && (abs(total[0]) != (2 << 4)) && (abs(total[1]) != (2 << 4)))
  {
    int32_t t[2] = {0, 0};
    uint64_t sb[5] = {p[0], p[-1], p[-5], p[1], p[5]};
    sub_pel_error_srfc(sb, t);
    total[0] += (int16_t)t[0];
    total[1] += (int16_t)t[1];
  }
  delta_mv[0] = total[0];
  delta_mv[1] = total[1];
  *min_cost = minCost;
  return notZero;
}

/* xDMVRCost alone (InterPrediction.cpp:2147-2155), exported so tests can pin it against oracle/_ref. */
uint64_t orc_dmvr_cost(const int16_t* a, int as, const int16_t* b, int bs, int w, int h, int simd_sad)
{
  return dmvr_cost(a, as, b, bs, w, h, simd_sad);
}

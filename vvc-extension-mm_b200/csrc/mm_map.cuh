// mm_map.cuh -- sphere / motion-model coordinate mapping of one 4x4 luma (2x2 chroma) sub-block.
//
// Product code (device + host-compilable).  Follows the reference's float operation order:
//   MVReprojection::reprojectMotionVectorSubblocks   source/Lib/CommonLib/MVReprojection.cpp:80-166
//   MotionPlaneAdaptiveMotionModel                   MotionModels/MotionPlaneAdaptiveMotionModel.cpp:26-187
//   GeodesicMotionModel                              MotionModels/GeodesicMotionModel.cpp:14-176
//   RotationalMotionModel                            MotionModels/RotationalMotionModel.cpp:8-78
//   TangentialMotionModel                            MotionModels/TangentialMotionModel.cpp:8-48
//   ThreeDTranslationalMotionModel                   MotionModels/ThreeDTranslationalMotionModel.cpp:7-24
//   Equirectangular/PerspectiveProjection            Projection.cpp:119-249,  Coordinate.cpp:8-68
// Every float operation goes through MM_F* (one IEEE rounding each, never contracted); libm and
// Eigen packet math come from mm_libm.cuh.  `pk` says whether the element sits in the vectorised
// body of Eigen's array traversal (linear index < (n/4)*4), SURVEY risk R1.
#pragma once
#include "mm_libm.cuh"

enum MmModel
{
  MM_CLASSIC = 0, MM_MPA_FRONT_BACK = 1, MM_MPA_LEFT_RIGHT = 2, MM_MPA_TOP_BOTTOM = 3, MM_TANGENTIAL = 4,
  MM_THREE_D_TRANSLATIONAL = 5, MM_ROTATIONAL = 6, MM_GEODESIC_X = 7, MM_GEODESIC_Y = 8, MM_GEODESIC_Z = 9,
  MM_GEODESIC_CAMPOSE = 10, MM_NUM_MODELS = 11
};

// Sequence-level constants (per context).
struct MmGeom
{
  float picW, picH;   // TCoord(m_resolution.width/height)
  float focal;        // TCoord(1./tan(M_PI/H))            Projection.h:128
  float angRes;       // TCoord(M_PI/H)                    MVReprojection.cpp:27,33,39
  float off;          // m_offset4x4
  int rows4, cols4;   // H/4, W/4
  int gedModulated;   // GeodesicMotionModel::VISHWANATH_MODULATED
  int fmaBuild;       // host glibc uses __sinf_fma/__cosf_fma
};

// Per (block, mv) scalars, computed once per candidate (mm_cand_setup) and shared by its sub-blocks.
struct MmCand
{
  float mvX, mvY;   // MVReprojection.cpp:123-124
  float p[9];       // ROT: rotationMatrixReally (row-major) | 3DT: d[0..2] | TAN: epsC? see below | GED: k
  int mvZero;
};
// TAN layout in p[]: p[0]=alphaC, p[1]=sinC, p[2]=cosC, p[3]=mvY*angRes, p[4]=mvX*angRes
// GED layout: p[0]=k, p[1]=angRes*mvX, p[2]=angRes*mvY

#define MM_PI_F 3.14159274101257324f /* TCoord(M_PI) */
#define MM_2PI_F 6.28318548202514648f /* TCoord(2)*TCoord(M_PI) */
#define MM_PI_2_F 1.57079637050628662f /* TCoord(M_PI_2) */

MM_HD float mm_asin_pk(float x, bool pk, bool fma) { return pk ? mm_psin(x) : mm_sinf(x, fma); }
MM_HD float mm_acos_pk(float x, bool pk, bool fma) { return pk ? mm_pcos(x) : mm_cosf(x, fma); }

// Array sin & cos of the same argument (packet: one shared Cephes core).
MM_HD void mm_sincos_pk(float x, bool pk, bool fma, float* s, float* c)
{
  if (pk)
  {
    int32_t j;
    float ys, yc;
    mm_ps_core(mm_fabsf(x), &j, &ys, &yc);
    const uint32_t sgn = MM_F2U(x) & 0x80000000u;
    const uint32_t swap = ((uint32_t)j & 4u) << 29;
    const bool ps = (((uint32_t)j & 2u) == 0u);
    *s = MM_U2F(MM_F2U(ps ? ys : yc) ^ (sgn ^ swap));
    const uint32_t j2 = (uint32_t)j - 2u;
    const uint32_t csgn = ((~j2) & 4u) << 29;
    const bool pc = ((j2 & 2u) == 0u);
    *c = MM_U2F(MM_F2U(pc ? ys : yc) ^ csgn);
  }
  else
  {
    *s = mm_sinf(x, fma);
    *c = mm_cosf(x, fma);
  }
}

// Eigen scalar_min_op/scalar_max_op = std::min/std::max with the array value first (NaN propagates).
MM_HD float mm_eig_clamp1(float a)
{
  a = (1.0f < a) ? 1.0f : a;
  a = (a < -1.0f) ? -1.0f : a;
  return a;
}
// std::min(TCoord(1), std::max(TCoord(-1), v)) of the scalar overload (NaN -> -1).
MM_HD float mm_std_clamp1(float v)
{
  const float m = (-1.0f < v) ? v : -1.0f;
  return (m < 1.0f) ? m : 1.0f;
}
MM_HD float mm_dot3(float a0, float b0, float a1, float b1, float a2, float b2)
{
  return MM_FADD(MM_FMUL(a0, b0), MM_FADD(MM_FMUL(a1, b1), MM_FMUL(a2, b2)));
}
MM_HD void mm_mat3_mul(const float* A, const float* B, float* C)
{
  float T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      T[i * 3 + j] = mm_dot3(A[i * 3 + 0], B[0 * 3 + j], A[i * 3 + 1], B[1 * 3 + j], A[i * 3 + 2], B[2 * 3 + j]);
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = T[i];
}

// Eigen::AngleAxis<float>::toRotationMatrix for UnitZ / UnitY axes (general formula evaluated literally).
MM_HD void mm_angle_axis(float angle, float ax, float ay, float az, bool fma, float* R)
{
  const float s = mm_sinf(angle, fma), c = mm_cosf(angle, fma);
  const float sx = MM_FMUL(s, ax), sy = MM_FMUL(s, ay), sz = MM_FMUL(s, az);
  const float omc = MM_FSUB(1.0f, c);
  const float cx = MM_FMUL(omc, ax), cy = MM_FMUL(omc, ay), cz = MM_FMUL(omc, az);
  float tmp;
  tmp = MM_FMUL(cx, ay); R[1] = MM_FSUB(tmp, sz); R[3] = MM_FADD(tmp, sz);
  tmp = MM_FMUL(cx, az); R[2] = MM_FADD(tmp, sy); R[6] = MM_FSUB(tmp, sy);
  tmp = MM_FMUL(cy, az); R[5] = MM_FSUB(tmp, sx); R[7] = MM_FADD(tmp, sx);
  R[0] = MM_FADD(MM_FMUL(cx, ax), c);
  R[4] = MM_FADD(MM_FMUL(cy, ay), c);
  R[8] = MM_FADD(MM_FMUL(cz, az), c);
}

// ---- Coordinate.cpp ------------------------------------------------------------------------------
MM_HD void mm_arr_cart_to_sph(float x, float y, float z, float* r, float* th, float* ph)
{
  const float rr = MM_FSQRT(MM_FADD(MM_FADD(MM_FMUL(x, x), MM_FMUL(y, y)), MM_FMUL(z, z)));
  *r = rr;
  *th = mm_acosf(mm_eig_clamp1(MM_FDIV(z, rr)));
  *ph = mm_atan2f(y, x);
}
MM_HD void mm_sc_cart_to_sph(const float* c, float* s)
{
  const float r = MM_FSQRT(MM_FADD(MM_FADD(MM_FMUL(c[0], c[0]), MM_FMUL(c[1], c[1])), MM_FMUL(c[2], c[2])));
  s[0] = r;
  s[1] = mm_acosf(mm_std_clamp1(MM_FDIV(c[2], r)));
  s[2] = mm_atan2f(c[1], c[0]);
}
MM_HD void mm_arr_sph_to_cart(float r, float th, float ph, bool pk, bool fma, float* o)
{
  float st, ct, sp, cp;
  mm_sincos_pk(th, pk, fma, &st, &ct);
  mm_sincos_pk(ph, pk, fma, &sp, &cp);
  const float rs = MM_FMUL(r, st);
  o[0] = MM_FMUL(rs, cp);
  o[1] = MM_FMUL(rs, sp);
  o[2] = MM_FMUL(r, ct);
}
MM_HD void mm_sc_sph_to_cart(float r, float th, float ph, bool fma, float* o)
{
  const float rs = MM_FMUL(r, mm_sinf(th, fma));
  o[0] = MM_FMUL(rs, mm_cosf(ph, fma));
  o[1] = MM_FMUL(rs, mm_sinf(ph, fma));
  o[2] = MM_FMUL(r, mm_cosf(th, fma));
}

// ---- Projection.cpp: equirectangular -------------------------------------------------------------
MM_HD void mm_erp_angles(const MmGeom& g, float X, float Y, float* th, float* ph)
{
  const float x = MM_FADD(X, 0.0f), y = MM_FADD(Y, 0.0f);
  *ph = MM_FMUL(MM_FMUL(mm_negf(MM_FDIV(x, g.picW)), 2.0f), MM_PI_F);
  *th = MM_FMUL(MM_FDIV(y, g.picH), MM_PI_F);
}
MM_HD void mm_erp_to_sphere_arr(const MmGeom& g, float X, float Y, bool pk, float* o)
{
  float th, ph;
  mm_erp_angles(g, X, Y, &th, &ph);
  mm_arr_sph_to_cart(1.0f, th, ph, pk, g.fmaBuild != 0, o);
}
MM_HD void mm_erp_to_sphere_sc(const MmGeom& g, float X, float Y, float* o)
{
  float th, ph;
  mm_erp_angles(g, X, Y, &th, &ph);
  mm_sc_sph_to_cart(1.0f, th, ph, g.fmaBuild != 0, o);
}
MM_HD void mm_erp_from_sphere_arr(const MmGeom& g, float x, float y, float z, float* X, float* Y)
{
  float r, th, ph;
  mm_arr_cart_to_sph(x, y, z, &r, &th, &ph);
  if (ph > 0.0f) ph = MM_FSUB(ph, MM_2PI_F);
  // (the reference subtracts a zero sample offset here: v - 0.0f == v for every v, omitted)
  *X = MM_FMUL(mm_negf(MM_FDIV(ph, MM_2PI_F)), g.picW);
  *Y = MM_FMUL(MM_FDIV(th, MM_PI_F), g.picH);
}

// ---- Projection.cpp: perspective (motion plane) ----------------------------------------------------
MM_HD void mm_pers_from_sphere_arr(const MmGeom& g, float sx, float sy, float sz, bool pk, float* X, float* Y, bool* vip)
{
  float r, th, ph;
  mm_arr_cart_to_sph(sy, mm_negf(sz), mm_negf(sx), &r, &th, &ph);
  const float polarR = MM_FMUL(g.focal, mm_tanf(th));
  float sp, cp;
  mm_sincos_pk(ph, pk, g.fmaBuild != 0, &sp, &cp);
  *X = MM_FADD(MM_FMUL(polarR, cp), 0.0f);
  *Y = MM_FADD(MM_FMUL(polarR, sp), 0.0f);
  *vip = polarR < 0.0f;
}
MM_HD void mm_pers_to_sphere_arr(const MmGeom& g, float X, float Y, bool vip, bool pk, float* o)
{
  const float x = X, y = Y;  // X - 0.0f, Y - 0.0f in the reference: identities
  const float polarR = MM_FSQRT(MM_FADD(MM_FMUL(x, x), MM_FMUL(y, y)));
  const float polarPhi = mm_atan2f(y, x);
  float th = mm_atanf(MM_FDIV(polarR, g.focal));
  const float vf = vip ? 1.0f : 0.0f;
  th = MM_FSUB(th, MM_FMUL(vf, MM_FSUB(MM_FMUL(2.0f, th), MM_PI_F)));
  const float ph = MM_FSUB(polarPhi, MM_FMUL(vf, MM_PI_F));
  float s[3];
  mm_arr_sph_to_cart(1.0f, th, ph, pk, g.fmaBuild != 0, s);
  o[0] = mm_negf(s[2]);
  o[1] = s[0];
  o[2] = mm_negf(s[1]);
}
// ERP toSphere of a grid position; `tr` (optional) = {sin(theta), cos(theta), sin(phi), cos(phi)} of that position
// taken from the per-picture tables (same functions, same values: the tables only hoist the evaluation).
struct MmTrig  // table values of one grid position; valid == false -> evaluate directly
{
  bool valid;
  float st, ct, sp, cp;
};
MM_HD void mm_erp_to_sphere_grid(const MmGeom& g, float X, float Y, bool pk, const MmTrig& tr, float* o)
{
  if (tr.valid)
  {
    o[0] = MM_FMUL(tr.st, tr.cp);  // (1*sin(theta))*cos(phi): 1*x is exact
    o[1] = MM_FMUL(tr.st, tr.sp);
    o[2] = tr.ct;
  }
  else
  {
    mm_erp_to_sphere_arr(g, X, Y, pk, o);
  }
}
MM_HD void mm_mpa_to_perspective(const MmGeom& g, int model, float X, float Y, bool pk, const MmTrig& tr, float* PX,
                                 float* PY, bool* vip)
{
  float s[3];
  mm_erp_to_sphere_grid(g, X, Y, pk, tr, s);
  float mx, my, mz;
  if (model == MM_MPA_FRONT_BACK) { mx = s[0]; my = s[1]; mz = s[2]; }
  else if (model == MM_MPA_LEFT_RIGHT) { mx = s[1]; my = mm_negf(s[0]); mz = s[2]; }
  else { mx = mm_negf(s[2]); my = s[1]; mz = s[0]; }
  mm_pers_from_sphere_arr(g, mx, my, mz, pk, PX, PY, vip);
}
// ---- per-candidate scalars -----------------------------------------------------------------------
// blk = position/size of pu.blocks[compID] (component units, quirk Q1); R = rotation of the GED model.
// mvX, mvY: the motion vector as float (MVReprojection.cpp:123-124 for the 1/16-pel MVs of MC / ME; the MM-MVP conversion
// passes MVs of other precisions, MVReprojection.cpp:186-187)
MM_HD void mm_cand_setup_float(const MmGeom& g, int model, bool isLuma, int posX, int posY, int sizeW, int sizeH, float mvX,
                               float mvY, const float* R, MmCand* c)
{
  const bool fma = g.fmaBuild != 0;
  c->mvX = mvX;
  c->mvY = mvY;
  c->mvZero = (mvX == 0.0f && mvY == 0.0f) ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) c->p[i] = 0.0f;
  const float bcx = MM_FADD((float)posX, MM_FDIV(MM_FSUB((float)sizeW, 1.0f), 2.0f));
  const float bcy = MM_FADD((float)posY, MM_FDIV(MM_FSUB((float)sizeH, 1.0f), 2.0f));
  if (model == MM_THREE_D_TRANSLATIONAL)
  {
    if (c->mvZero) return;
    float a[3], b[3];
    mm_erp_to_sphere_sc(g, bcx, bcy, a);
    mm_erp_to_sphere_sc(g, MM_FADD(bcx, mvX), MM_FADD(bcy, mvY), b);
    c->p[0] = MM_FSUB(b[0], a[0]);
    c->p[1] = MM_FSUB(b[1], a[1]);
    c->p[2] = MM_FSUB(b[2], a[2]);
  }
  else if (model == MM_ROTATIONAL)
  {
    if (c->mvZero) return;
    float cc[3], sc[3], Rz[9], Ry[9], rot[9], uphi[9], uth[9], unrot[9], unrotT[9], t[9];
    mm_erp_to_sphere_sc(g, bcx, bcy, cc);
    mm_sc_cart_to_sph(cc, sc);
    mm_angle_axis(MM_FMUL(mm_negf(mvX), g.angRes), 0.f, 0.f, 1.f, fma, Rz);
    mm_angle_axis(MM_FMUL(mvY, g.angRes), 0.f, 1.f, 0.f, fma, Ry);
    mm_mat3_mul(Rz, Ry, rot);
    mm_angle_axis(mm_negf(sc[2]), 0.f, 0.f, 1.f, fma, uphi);
    mm_angle_axis(MM_D2F(MM_DSUB(0x1.921fb54442d18p+0, (double)sc[1])), 0.f, 1.f, 0.f, fma, uth);
    mm_mat3_mul(uth, uphi, unrot);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) unrotT[i * 3 + j] = unrot[j * 3 + i];
    mm_mat3_mul(rot, unrot, t);
    mm_mat3_mul(unrotT, t, c->p);
  }
  else if (model == MM_TANGENTIAL)
  {
    // block constants (block centre only): filled for zero MVs too, the candidate loop reads them from any candidate
    float cc[3], sc[3];
    mm_erp_to_sphere_sc(g, bcx, bcy, cc);
    mm_sc_cart_to_sph(cc, sc);
    const float epsC = MM_D2F(MM_DSUB(0x1.921fb54442d18p+0, (double)sc[1]));
    c->p[0] = sc[2];
    // unqualified sin()/cos() on a float bind to the double libm overloads in the reference (R4): narrowed glibc
    // results, reproduced exactly on the device (mm_libm.cuh: mm_sin_d2f / mm_cos_d2f)
    c->p[1] = mm_sin_d2f(epsC);
    c->p[2] = mm_cos_d2f(epsC);
    c->p[3] = MM_FMUL(mvY, g.angRes);
    c->p[4] = MM_FMUL(mvX, g.angRes);
  }
  else if (model >= MM_GEODESIC_X && model <= MM_GEODESIC_CAMPOSE)
  {
    c->p[1] = MM_FMUL(g.angRes, mvX);
    c->p[2] = MM_FMUL(g.angRes, mvY);
    if (g.gedModulated && (isLuma || !c->mvZero))
    {
      float cc[3], cr[3], sc[3];
      mm_erp_to_sphere_sc(g, bcx, bcy, cc);
      cr[0] = mm_dot3(R[0], cc[0], R[1], cc[1], R[2], cc[2]);
      cr[1] = mm_dot3(R[3], cc[0], R[4], cc[1], R[5], cc[2]);
      cr[2] = mm_dot3(R[6], cc[0], R[7], cc[1], R[8], cc[2]);
      mm_sc_cart_to_sph(cr, sc);
      c->p[0] = MM_FDIV(mm_sinf(MM_FADD(sc[1], c->p[1]), fma), mm_sinf(c->p[1], fma));
    }
  }
}

MM_HD void mm_cand_setup(const MmGeom& g, int model, bool isLuma, int posX, int posY, int sizeW, int sizeH, int mvHor,
                         int mvVer, const float* R, MmCand* c)
{
  const float mvX = MM_FADD((float)(mvHor >> 4), MM_FDIV((float)(mvHor & 15), 16.0f));
  const float mvY = MM_FADD((float)(mvVer >> 4), MM_FDIV((float)(mvVer & 15), 16.0f));
  mm_cand_setup_float(g, model, isLuma, posX, posY, sizeW, sizeH, mvX, mvY, R, c);
}

// ---- one sub-block -------------------------------------------------------------------------------
// The mapping of a sub-block is split into the part that does not depend on the candidate MV (mm_model_prepare:
// the reference recomputes it per call, encoder-side memoisation aside -- quirk Q8 says that has no numeric effect)
// and the per-candidate part (mm_model_apply).  k_map prepares once per chunk of sub-blocks and applies per candidate.
struct MmInv
{
  float a, b, c, d, e;  // MPA: PX,PY | GED: r,theta,phi,sinf(theta),cosf(theta) | ROT/3DT: s0,s1,s2 | TAN: yPl,xPl
  bool flag;            // MPA: virtual image plane
};

// X,Y: luma-scale coordinates of the sub-block (4*i + off).  For luma MPA the caller passes the cached motion-plane
// coordinates (cPX,cPY,cVip) of the picture-wide cache.  `tr` = optional table values of the original position.
// `blk` = any candidate of the block (TAN block constants p[0..2] are MV independent).
MM_HD void mm_model_prepare(const MmGeom& g, int model, bool isLuma, const MmCand& blk, const float* R, float X, float Y,
                            bool pk, float cPX, float cPY, bool cVip, const MmTrig& tr, MmInv* inv)
{
  const bool fma = g.fmaBuild != 0;
  inv->a = inv->b = inv->c = inv->d = inv->e = 0.0f;
  inv->flag = false;
  if (model >= MM_MPA_FRONT_BACK && model <= MM_MPA_TOP_BOTTOM)
  {
    float PX = cPX, PY = cPY;
    bool vip = cVip;
    if (!isLuma) mm_mpa_to_perspective(g, model, X, Y, pk, tr, &PX, &PY, &vip);
    inv->a = PX; inv->b = PY; inv->flag = vip;
    return;
  }
  float s[3];
  mm_erp_to_sphere_grid(g, X, Y, pk, tr, s);
  if (model >= MM_GEODESIC_X && model <= MM_GEODESIC_CAMPOSE)
  {
    float r, th, ph;
    const float rx = mm_dot3(R[0], s[0], R[1], s[1], R[2], s[2]);
    const float ry = mm_dot3(R[3], s[0], R[4], s[1], R[5], s[2]);
    const float rz = mm_dot3(R[6], s[0], R[7], s[1], R[8], s[2]);
    mm_arr_cart_to_sph(rx, ry, rz, &r, &th, &ph);
    inv->a = r; inv->b = th; inv->c = ph;
    if (g.gedModulated) { inv->d = mm_sinf(th, fma); inv->e = mm_cosf(th, fma); }
  }
  else if (model == MM_TANGENTIAL)
  {
    const float alphaC = blk.p[0], sinC = blk.p[1], cosC = blk.p[2];
    float r, th, ph;
    mm_arr_cart_to_sph(s[0], s[1], s[2], &r, &th, &ph);
    const float eps = MM_FSUB(MM_PI_2_F, th);
    const float dAl = MM_FSUB(ph, alphaC);
    float se, ce, sd, cd;
    mm_sincos_pk(eps, pk, fma, &se, &ce);
    mm_sincos_pk(dAl, pk, fma, &sd, &cd);
    const float cosPsi = MM_FADD(MM_FMUL(sinC, se), MM_FMUL(MM_FMUL(cosC, ce), cd));
    inv->a = MM_FDIV(MM_FSUB(MM_FMUL(se, cosC), MM_FMUL(MM_FMUL(sinC, ce), cd)), cosPsi);  // cart2DYPlane
    inv->b = MM_FDIV(MM_FMUL(sd, ce), cosPsi);                                              // cart2DXPlane
  }
  else
  {
    inv->a = s[0]; inv->b = s[1]; inv->c = s[2];
  }
}

// Returns moved coordinates (float, before NaN handling).  Every model ends in EquirectangularProjection::fromSphere
// of the moved point u: one shared tail.
MM_HD void mm_model_apply(const MmGeom& g, int model, bool isLuma, const MmCand& c, const float* R, float X, float Y,
                          bool pk, const MmInv& inv, float* MX, float* MY)
{
  const bool fma = g.fmaBuild != 0;
  *MX = X;
  *MY = Y;
  float u0, u1, u2;
  if (model >= MM_MPA_FRONT_BACK && model <= MM_MPA_TOP_BOTTOM)
  {
    const bool vip = inv.flag;
    const float sgn = vip ? -1.0f : 1.0f;
    const float QX = MM_FADD(inv.a, MM_FMUL(c.mvX, sgn));
    const float QY = MM_FADD(inv.b, MM_FMUL(c.mvY, sgn));
    float m[3];
    mm_pers_to_sphere_arr(g, QX, QY, vip, pk, m);  // MotionPlaneAdaptiveMotionModel::toProjection (:163-187)
    if (model == MM_MPA_FRONT_BACK) { u0 = m[0]; u1 = m[1]; u2 = m[2]; }
    else if (model == MM_MPA_LEFT_RIGHT) { u0 = mm_negf(m[1]); u1 = m[0]; u2 = m[2]; }
    else { u0 = m[2]; u1 = m[1]; u2 = mm_negf(m[0]); }
  }
  else
  {
    const bool ged = model >= MM_GEODESIC_X && model <= MM_GEODESIC_CAMPOSE;
    if (c.mvZero && !(ged && isLuma)) return;  // quirk Q3: identity for mv == 0 (not for cached GED)
    if (ged)
    {
      const float r = inv.a, th = inv.b, ph = inv.c;
      float thM;
      if (!g.gedModulated) thM = MM_FADD(th, c.p[1]);
      else thM = MM_FADD(th, mm_atanf(MM_FDIV(inv.d, MM_FSUB(c.p[0], inv.e))));
      const float phM = MM_FADD(ph, c.p[2]);
      float q[3];
      mm_arr_sph_to_cart(r, thM, phM, pk, fma, q);
      u0 = mm_dot3(R[0], q[0], R[3], q[1], R[6], q[2]);
      u1 = mm_dot3(R[1], q[0], R[4], q[1], R[7], q[2]);
      u2 = mm_dot3(R[2], q[0], R[5], q[1], R[8], q[2]);
    }
    else if (model == MM_ROTATIONAL)
    {
      u0 = mm_dot3(c.p[0], inv.a, c.p[1], inv.b, c.p[2], inv.c);
      u1 = mm_dot3(c.p[3], inv.a, c.p[4], inv.b, c.p[5], inv.c);
      u2 = mm_dot3(c.p[6], inv.a, c.p[7], inv.b, c.p[8], inv.c);
    }
    else if (model == MM_THREE_D_TRANSLATIONAL)
    {
      u0 = MM_FADD(inv.a, c.p[0]);
      u1 = MM_FADD(inv.b, c.p[1]);
      u2 = MM_FADD(inv.c, c.p[2]);
    }
    else if (model == MM_TANGENTIAL)
    {
      const float alphaC = c.p[0], sinC = c.p[1], cosC = c.p[2];
      const float yM = MM_FSUB(inv.a, c.p[3]);
      const float xM = MM_FSUB(inv.b, c.p[4]);
      const float rho = MM_FSQRT(MM_FADD(MM_FMUL(xM, xM), MM_FMUL(yM, yM)));
      const float eta = mm_atanf(rho);
      float sEp, cEp;
      mm_sincos_pk(eta, pk, fma, &sEp, &cEp);
      const float gamma = MM_FSUB(MM_FMUL(MM_FMUL(rho, cosC), cEp), MM_FMUL(MM_FMUL(yM, sinC), sEp));
      // non-vectorised expressions (atan / asin): scalar libm sin/cos
      float sEs = sEp, cEs = cEp;
      if (pk) { sEs = mm_sinf(eta, fma); cEs = mm_cosf(eta, fma); }
      const float alphaM = MM_FADD(alphaC, mm_atanf(MM_FDIV(MM_FMUL(xM, sEs), gamma)));
      const float epsM = mm_asinf(MM_FADD(MM_FMUL(cEs, sinC), MM_FDIV(MM_FMUL(MM_FMUL(yM, sEs), cosC), rho)));
      float q[3];
      mm_arr_sph_to_cart(1.0f, MM_FSUB(MM_PI_2_F, epsM), alphaM, pk, fma, q);
      u0 = q[0]; u1 = q[1]; u2 = q[2];
    }
    else
    {
      return;
    }
  }
  mm_erp_from_sphere_arr(g, u0, u1, u2, MX, MY);
}

MM_HD void mm_model_motion(const MmGeom& g, int model, bool isLuma, const MmCand& c, const float* R, float X, float Y,
                           bool pk, float cPX, float cPY, bool cVip, const MmTrig& tr, float* MX, float* MY)
{
  MmInv inv;
  mm_model_prepare(g, model, isLuma, c, R, X, Y, pk, cPX, cPY, cVip, tr, &inv);
  mm_model_apply(g, model, isLuma, c, R, X, Y, pk, inv, MX, MY);
}

// MVReprojection.cpp:152-165: NaN -> no motion, remove offset, chroma rescale, std::round to fixed point.
MM_HD void mm_to_fixed(const MmGeom& g, float X, float Y, float MX, float MY, int csx, int csy, int* fx, int* fy)
{
  const bool nan = mm_isnanf(MX) || mm_isnanf(MY);
  float x = MM_FSUB(nan ? X : MX, g.off), y = MM_FSUB(nan ? Y : MY, g.off);
  if (csx | csy)
  {
    x = MM_FDIV(x, (float)(1 << csx));
    y = MM_FDIV(y, (float)(1 << csy));
  }
  *fx = mm_cvtt_f32_i32(mm_roundf(MM_FMUL(x, (float)(16 << csx))));
  *fy = mm_cvtt_f32_i32(mm_roundf(MM_FMUL(y, (float)(16 << csy))));
}

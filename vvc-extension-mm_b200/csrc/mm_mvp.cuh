// mm_mvp.cuh -- MM-MVP conversion: MVReprojection::motionVectorInDesiredMotionModel
// (CommonLib/MVReprojection.cpp:168-217) and every model's scalar motionVectorForEquivalentPixelShiftAt
//   MotionPlaneAdaptiveMotionModel.cpp:76-100, 137-162      GeodesicMotionModel.cpp:178-221
//   RotationalMotionModel.cpp:80-100                        TangentialMotionModel.cpp:50-83
//   ThreeDTranslationalMotionModel.cpp:26-38                TranslationalMotionModel.cpp:14-17
// The reference evaluates it on the host, one candidate at a time (UnitTools.cpp:2272,2302,2932,...): single points, so
// every Eigen expression is below the packet width and runs the scalar std:: functions.  Same one-operation-per-statement
// style as mm_map.cuh; compiled for the host (mmgpu_mvp_convert) and for the device.
#pragma once
#include "mm_map.cuh"

// EquirectangularProjection::fromSphere(const Array3TCoord&), Projection.cpp:241-249 (scalar cartesianToSpherical)
MM_HD void mm_erp_from_sphere_sc(const MmGeom& g, const float* c, float* X, float* Y)
{
  float s[3];
  mm_sc_cart_to_sph(c, s);
  float ph = s[2];
  if (ph > 0.0f) ph = MM_FSUB(ph, MM_2PI_F);
  *X = MM_FMUL(mm_negf(MM_FDIV(ph, MM_2PI_F)), g.picW);   // "- m_pixelOffset" with offset 0: identity
  *Y = MM_FMUL(MM_FDIV(s[1], MM_PI_F), g.picH);
}

// MotionPlaneAdaptiveMotionModel::toPerspective(const Array2TCoord&) (:137-162) -> PerspectiveProjection::fromSphere(const
// Array3TCoord&) (Projection.cpp:181-196): scalar libm throughout
MM_HD void mm_mpa_to_perspective_sc(const MmGeom& g, int model, float X, float Y, float* PX, float* PY, bool* vip)
{
  const bool fma = g.fmaBuild != 0;
  float s[3];
  mm_erp_to_sphere_sc(g, X, Y, s);
  float mx, my, mz;
  if (model == MM_MPA_FRONT_BACK) { mx = s[0]; my = s[1]; mz = s[2]; }
  else if (model == MM_MPA_LEFT_RIGHT) { mx = s[1]; my = mm_negf(s[0]); mz = s[2]; }
  else { mx = mm_negf(s[2]); my = s[1]; mz = s[0]; }
  const float rot[3] = {my, mm_negf(mz), mm_negf(mx)};
  float sp[3];
  mm_sc_cart_to_sph(rot, sp);
  const float polarR = MM_FMUL(g.focal, mm_tanf(sp[1]));
  *PX = MM_FADD(MM_FMUL(polarR, mm_cosf(sp[2], fma)), 0.0f);
  *PY = MM_FADD(MM_FMUL(polarR, mm_sinf(sp[2], fma)), 0.0f);
  *vip = polarR < 0.0f;
}

// R * v for a 3-vector (coefficient-based product: a0*b0 + (a1*b1 + a2*b2))
MM_HD void mm_mat3_vec(const float* R, const float* v, float* o)
{
  o[0] = mm_dot3(R[0], v[0], R[1], v[1], R[2], v[2]);
  o[1] = mm_dot3(R[3], v[0], R[4], v[1], R[5], v[2]);
  o[2] = mm_dot3(R[6], v[0], R[7], v[1], R[8], v[2]);
}

// motionVectorForEquivalentPixelShiftAt of `model`: position (integer sample), shifted position, block centre of the CURRENT block
MM_HD void mm_mvp_equivalent(const MmGeom& g, int model, const float* R, int posX, int posY, float shX, float shY, float bcx, float bcy,
                             float* mvx, float* mvy)
{
  const bool fma = g.fmaBuild != 0;
  const float px = (float)posX, py = (float)posY;
  if (model == MM_CLASSIC)
  {
    *mvx = MM_FSUB(shX, px);
    *mvy = MM_FSUB(shY, py);
  }
  else if (model >= MM_MPA_FRONT_BACK && model <= MM_MPA_TOP_BOTTOM)
  {
    float ox, oy, sx, sy;
    bool vo, vs;
    mm_mpa_to_perspective_sc(g, model, px, py, &ox, &oy, &vo);
    mm_mpa_to_perspective_sc(g, model, shX, shY, &sx, &sy, &vs);
    if (vo != vs) { *mvx = 0.0f; *mvy = 0.0f; return; }
    const float sgn = vs ? -1.0f : 1.0f;
    *mvx = MM_FMUL(MM_FSUB(sx, ox), sgn);
    *mvy = MM_FMUL(MM_FSUB(sy, oy), sgn);
  }
  else if (model == MM_THREE_D_TRANSLATIONAL)
  {
    float cc[3], c[3], cm[3], t[3];
    mm_erp_to_sphere_sc(g, bcx, bcy, cc);
    mm_erp_to_sphere_sc(g, px, py, c);
    mm_erp_to_sphere_sc(g, shX, shY, cm);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = MM_FADD(MM_FSUB(cm[i], c[i]), cc[i]);
    float X, Y;
    mm_erp_from_sphere_sc(g, t, &X, &Y);
    *mvx = MM_FSUB(X, bcx);
    *mvy = MM_FSUB(Y, bcy);
  }
  else if (model == MM_ROTATIONAL)
  {
    float cc[3], sc[3], c[3], cm[3], uphi[9], uth[9], U[9], a[3], am[3], sa[3], sm[3];
    mm_erp_to_sphere_sc(g, bcx, bcy, cc);
    mm_sc_cart_to_sph(cc, sc);
    mm_erp_to_sphere_sc(g, px, py, c);
    mm_erp_to_sphere_sc(g, shX, shY, cm);
    mm_angle_axis(mm_negf(sc[2]), 0.f, 0.f, 1.f, fma, uphi);
    mm_angle_axis(MM_D2F(MM_DSUB(0x1.921fb54442d18p+0, (double)sc[1])), 0.f, 1.f, 0.f, fma, uth);
    mm_mat3_mul(uth, uphi, U);
    mm_mat3_vec(U, c, a);
    mm_mat3_vec(U, cm, am);
    mm_sc_cart_to_sph(a, sa);
    mm_sc_cart_to_sph(am, sm);
    *mvx = MM_FDIV(MM_FSUB(sa[2], sm[2]), g.angRes);
    *mvy = MM_FDIV(MM_FSUB(sm[1], sa[1]), g.angRes);
  }
  else if (model == MM_TANGENTIAL)
  {
    float cc[3], sc[3], c[3], s[3], cm[3], sm[3];
    mm_erp_to_sphere_sc(g, bcx, bcy, cc);
    mm_sc_cart_to_sph(cc, sc);
    const float eC = MM_FSUB(MM_PI_2_F, sc[1]), aC = sc[2];
    mm_erp_to_sphere_sc(g, px, py, c);
    mm_sc_cart_to_sph(c, s);
    const float e = MM_FSUB(MM_PI_2_F, s[1]), al = s[2];
    mm_erp_to_sphere_sc(g, shX, shY, cm);
    mm_sc_cart_to_sph(cm, sm);
    const float eM = MM_FSUB(MM_PI_2_F, sm[1]), alM = sm[2];
    const float sEC = mm_sinf(eC, fma), cEC = mm_cosf(eC, fma);
    float yP[2], xP[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
    {
      const float ee = k ? eM : e, dA = MM_FSUB(k ? alM : al, aC);
      const float se = mm_sinf(ee, fma), ce = mm_cosf(ee, fma), cd = mm_cosf(dA, fma), sd = mm_sinf(dA, fma);
      const float cosPsi = MM_FADD(MM_FMUL(sEC, se), MM_FMUL(MM_FMUL(cEC, ce), cd));
      yP[k] = MM_FDIV(MM_FSUB(MM_FMUL(se, cEC), MM_FMUL(MM_FMUL(sEC, ce), cd)), cosPsi);
      xP[k] = MM_FDIV(MM_FMUL(sd, ce), cosPsi);
    }
    *mvx = MM_FDIV(MM_FSUB(xP[0], xP[1]), g.angRes);
    *mvy = MM_FDIV(MM_FSUB(yP[0], yP[1]), g.angRes);
  }
  else   // geodesic
  {
    float c[3], cr[3], s[3], cm[3], cmr[3], sm[3];
    mm_erp_to_sphere_sc(g, px, py, c);
    mm_mat3_vec(R, c, cr);
    mm_sc_cart_to_sph(cr, s);
    mm_erp_to_sphere_sc(g, shX, shY, cm);
    mm_mat3_vec(R, cm, cmr);
    mm_sc_cart_to_sph(cmr, sm);
    if (!g.gedModulated)
    {
      *mvx = MM_FDIV(MM_FSUB(sm[1], s[1]), g.angRes);
      *mvy = MM_FDIV(MM_FSUB(sm[2], s[2]), g.angRes);
    }
    else
    {
      float cc[3], ccr[3], sc[3];
      mm_erp_to_sphere_sc(g, bcx, bcy, cc);
      mm_mat3_vec(R, cc, ccr);
      mm_sc_cart_to_sph(ccr, sc);
      const float dTh = MM_FSUB(sm[1], s[1]);
      const float k = MM_FDIV(mm_sinf(MM_FADD(dTh, s[1]), fma), mm_sinf(dTh, fma));
      const float dThC = mm_atanf(MM_FDIV(mm_sinf(sc[1], fma), MM_FSUB(k, mm_cosf(sc[1], fma))));
      *mvx = MM_FDIV(dThC, g.angRes);
      *mvy = MM_FDIV(MM_FSUB(sm[2], s[2]), g.angRes);
    }
  }
}

// MVReprojection::motionVectorInDesiredMotionModel.  Rorig / Rdes: rotation of the geodesic models (any pointer otherwise);
// sameRotation: the two (curPOC, refPOC) pairs have the same epipole (only read for GEODESIC_CAMPOSE -> GEODESIC_CAMPOSE).
MM_HD void mm_mvp_convert(const MmGeom& g, int posX, int posY, int mvHor, int mvVer, int modelOrig, int modelDes, int shiftHor,
                          int shiftVer, const float* Rorig, const float* Rdes, bool sameRotation, int candX, int candY, int candW,
                          int candH, int curX, int curY, int curW, int curH, int* outHor, int* outVer)
{
  *outHor = 0; *outVer = 0;
  if (mvHor == 0 && mvVer == 0) return;
  if (modelDes == modelOrig && (modelDes != MM_GEODESIC_CAMPOSE || sameRotation)) { *outHor = mvHor; *outVer = mvVer; return; }
  const float mvX = MM_FADD((float)(mvHor >> shiftHor), MM_FDIV((float)(mvHor & ((1 << shiftHor) - 1)), (float)(1 << shiftHor)));
  const float mvY = MM_FADD((float)(mvVer >> shiftVer), MM_FDIV((float)(mvVer & ((1 << shiftVer) - 1)), (float)(1 << shiftVer)));
  const float px = (float)posX, py = (float)posY;
  float shX, shY;
  if (modelOrig == MM_CLASSIC)
  {
    shX = MM_FADD(px, mvX);   // TranslationalMotionModel::modelMotion
    shY = MM_FADD(py, mvY);
  }
  else
  {
    // modelMotion (the un-cached variant, a 1x1 array: scalar traversal) of the single point `position`, block = candidate block
    MmCand c2;
    mm_cand_setup_float(g, modelOrig, false, candX, candY, candW, candH, mvX, mvY, Rorig, &c2);
    MmTrig none; none.valid = false; none.st = none.ct = none.sp = none.cp = 0.f;
    mm_model_motion(g, modelOrig, false, c2, Rorig, px, py, false, 0.f, 0.f, false, none, &shX, &shY);
  }
  const float bcx = MM_FADD((float)curX, MM_FDIV(MM_FSUB((float)curW, 1.0f), 2.0f));
  const float bcy = MM_FADD((float)curY, MM_FDIV(MM_FSUB((float)curH, 1.0f), 2.0f));
  float dx, dy;
  mm_mvp_equivalent(g, modelDes, Rdes, posX, posY, shX, shY, bcx, bcy, &dx, &dy);
  if (mm_isnanf(dx) || mm_isnanf(dy)) return;
  *outHor = mm_cvtt_f32_i32(mm_roundf(MM_FMUL(dx, (float)(1 << shiftHor))));
  *outVer = mm_cvtt_f32_i32(mm_roundf(MM_FMUL(dy, (float)(1 << shiftVer))));
}

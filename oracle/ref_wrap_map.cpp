// ref_wrap_map.cpp -- C entry points over the REFERENCE's own float-mapping source (TEST INFRASTRUCTURE ONLY).
//
// oracle/Makefile compiles, from where they lie under /root/reference/source/Lib/CommonLib, the untouched
//   Coordinate.cpp, Projection.cpp, MVReprojection.cpp, MMConfig.cpp, MotionModels/{MotionPlaneAdaptive,Geodesic,
//   Rotational,Tangential,ThreeDTranslational,Translational}MotionModel.cpp
// against oracle/eigen_func_shim (a functional stand-in for the Eigen 3.3.7 subset they use: Eigen itself is not on
// the box) and links them with this file into oracle/_ref/libvtmmap.so.  It checks the oracle's restatement of the
// reference's operation order / model algebra against the real source; what stays unpinned is only the behaviour the
// stand-in asserts about Eigen (listed at the top of eigen_func_shim/Eigen/Dense).
#include <cstdlib>
#include <cstring>
#include "CommonLib/MVReprojection.h"

// ---- units of the reference that are NOT compiled: LUT variants (unused: toProjectionLUT is commented out,
//      MotionPlaneAdaptiveMotionModel.cpp:41,72), calibrated-lens tables, the POC->epipole map -------------------
LookupTable::LookupTable(std::function<TCoord(TCoord)>, std::pair<TCoord, TCoord>, int) {}
TCoord LookupTable::lookup(TCoord) const { abort(); }
ArrayXXTCoordPtr LookupTable::lookup(ArrayXXTCoordPtr) const { abort(); }
TCoord LookupTable::inverseLookup(TCoord) const { abort(); }
ArrayXXTCoordPtr LookupTable::inverseLookup(ArrayXXTCoordPtr) const { abort(); }
void ReprojectionLUT::fill() {}  // two 2786x2728 tables the path never reads
Array2TCoord ReprojectionLUT::operator()(const Array2TCoord&) const { abort(); }
static Array3TCoord g_epipole(0, 0, 0), g_epipole2(0, 0, 0);   // the second one answers for current POC 7 (MM-MVP tests)
void EpipoleList::addEpipole(const Array3TCoord&, const int, const int, bool) {}
Array3TCoord EpipoleList::findEpipole(int curPOC, int) const { return curPOC == 7 ? g_epipole2 : g_epipole; }

// MVReprojection::init needs an SPS (Slice.cpp and half of CommonLib): set the same members directly.
struct RefMap : public MVReprojection
{
  EquirectangularProjection* proj;
  EpipoleList epl;
  RefMap(int w, int h, float off, int gedFlavor)
  {
    proj = new EquirectangularProjection(Size(w, h));
    m_projection = proj;
    m_resolution = Size(w, h);
    m_offset4x4 = off;
    m_epipoleList = &epl;
    for (auto& m : m_motionModels) m = nullptr;
    fillCache();                                                    // MVReprojection.cpp:67-71
    const double ang = M_PI / h;                                    // :27,33,39
    for (int id = MPA_FRONT_BACK; id <= MPA_TOP_BOTTOM; id++)       // :19-24
    {
      auto* mm = new MotionPlaneAdaptiveMotionModel(proj, MotionModelID(id));
      mm->fillCache({m_cart2DProj[0], m_cart2DProj[1]});
      m_motionModels[id] = mm;
    }
    m_motionModels[CLASSIC] = new TranslationalMotionModel();      // :16-18
    m_motionModels[TANGENTIAL] = new TangentialMotionModel(proj, ang);
    m_motionModels[THREE_D_TRANSLATIONAL] = new ThreeDTranslationalMotionModel(proj);
    m_motionModels[ROTATIONAL] = new RotationalMotionModel(proj, ang);
    for (int id = GEODESIC_X; id <= GEODESIC_CAMPOSE; id++)         // :34-55
    {
      auto* g = new GeodesicMotionModel(proj, ang, GeodesicMotionModel::Flavor(gedFlavor));
      g->fillCache({m_cart2DProj[0], m_cart2DProj[1]});
      if (id == GEODESIC_X) g->setEpipole({1.0, 0.0, 0.0});
      if (id == GEODESIC_Y) g->setEpipole({0.0, 1.0, 0.0});
      if (id == GEODESIC_Z) g->setEpipole({0.0, 0.0, 1.0});
      m_motionModels[id] = g;
    }
    m_initialized = true;
  }
};

extern "C" void* refmap_create(int w, int h, float off, int gedFlavor) { return new RefMap(w, h, off, gedFlavor); }
extern "C" void refmap_set_epipole(float x, float y, float z) { g_epipole = Array3TCoord(x, y, z); g_epipole2 = g_epipole; }
extern "C" void refmap_set_epipole2(float x, float y, float z) { g_epipole2 = Array3TCoord(x, y, z); }
// MVReprojection::motionVectorInDesiredMotionModel (MVReprojection.cpp:168-217)
extern "C" void refmap_mvp(void* h, int px, int py, int mvh, int mvv, int modelOrig, int modelDes, int shiftHor, int shiftVer, int curPOCOrig,
                           int curPOCDes, int candX, int candY, int candW, int candH, int curX, int curY, int curW, int curH, int* out)
{
  RefMap* r = (RefMap*)h;
  const Mv m = r->motionVectorInDesiredMotionModel(Position(px, py), Mv(mvh, mvv), MotionModelID(modelOrig), MotionModelID(modelDes), shiftHor,
                                                    shiftVer, curPOCOrig, 0, curPOCDes, 0, Position(candX, candY), Size(candW, candH),
                                                    Position(curX, curY), Size(curW, curH));
  out[0] = m.hor; out[1] = m.ver;
}
// GeodesicMotionModel::setEpipole -> rotation matrix (row-major), through a throw-away model instance
extern "C" void refmap_rodrigues(void* h, float x, float y, float z, float* R)
{
  struct G : GeodesicMotionModel
  {
    G(const Projection* p) : GeodesicMotionModel(p, 0.0f, GeodesicMotionModel::VISHWANATH_MODULATED) {}
    void get(float* R) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i * 3 + j] = m_rotationMatrix(i, j); }
  };
  G g(((RefMap*)h)->proj);
  g.setEpipole(Array3TCoord(2.0f, 3.0f, 5.0f));  // make sure the "unchanged epipole" shortcut does not fire
  g.setEpipole(Array3TCoord(x, y, z));
  g.get(R);
}
extern "C" void refmap_reproject(void* h, int px, int py, int w, int hh, int mvh, int mvv, int model, int comp, int chfmt,
                                 int curPOC, int refPOC, int* outx, int* outy)
{
  RefMap* r = (RefMap*)h;
  ArrayXXFixedPtrPair p = r->reprojectMotionVectorSubblocks(Position(px, py), Size(w, hh), Mv(mvh, mvv), MotionModelID(model),
                                                            ComponentID(comp), ChromaFormat(chfmt), curPOC, refPOC);
  const ArrayXXFixed& X = *std::get<0>(p);
  const ArrayXXFixed& Y = *std::get<1>(p);
  for (Eigen::Index i = 0; i < X.size(); i++) { outx[i] = X.coeff(i); outy[i] = Y.coeff(i); }
}
// MPA perspective cache of one plane model, column-major like the reference's arrays
extern "C" void refmap_mpa_cache(void* h, int model, float* px, float* py, unsigned char* vip)
{
  struct M : MotionPlaneAdaptiveMotionModel
  {
    void get(float* px, float* py, unsigned char* vip)
    {
      for (Eigen::Index i = 0; i < m_cart2DPers[0]->size(); i++)
      {
        px[i] = m_cart2DPers[0]->coeff(i); py[i] = m_cart2DPers[1]->coeff(i); vip[i] = m_vip->coeff(i) ? 1 : 0;
      }
    }
  };
  static_cast<M*>(static_cast<MotionPlaneAdaptiveMotionModel*>(((RefMap*)h)->getMotionModel(MotionModelID(model))))->get(px, py, vip);
}

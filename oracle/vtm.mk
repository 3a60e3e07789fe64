# oracle/_ref/vtm: the WHOLE reference (CommonLib, EncoderLib, DecoderLib, Utilities, libmd5, EncoderApp, DecoderApp),
# compiled unchanged from where it lies under /root/reference against the functional Eigen stand-in
# (eigen_func_shim/; Eigen 3.3.7 is an un-vendored ExternalProject and there is no network).  TEST INFRASTRUCTURE.
#
#   make -f vtm.mk -j8 stock      ->  _ref/vtm/EncoderApp, _ref/vtm/DecoderApp          (stock CPU reference)
#   make -f vtm.mk -j8 gpu        ->  _ref/vtm/EncoderApp_gpu, _ref/vtm/DecoderApp_gpu  (patches/mmgpu.patch applied to a
#                                      build-time COPY of the touched units under _ref/vtm/patched/, linked with libmmgpu.so;
#                                      every other object is shared with the stock build)
#
# Flags follow the reference's release build (CMakeLists.txt:20-22,91-93; CommonLib/CMakeLists.txt:81-93): -O3 -msse4.1,
# per-directory -mavx/-mavx2/-msse4.2 with the USE_* defines; -include cstdint for gcc 13 (TypeDef.h:276), -Werror dropped.
REF    ?= /root/reference
SRC    := $(REF)/source
LIB    := $(SRC)/Lib
OUT    := _ref/vtm
O      := $(OUT)/obj
CXX    ?= g++
INC    := -Ieigen_func_shim -I$(LIB) -I$(LIB)/CommonLib -I$(LIB)/CommonLib/x86 -I$(LIB)/libmd5 -I$(LIB)/EncoderLib \
          -I$(LIB)/DecoderLib -I$(LIB)/Utilities
CXXF   := -std=c++11 -O3 -msse4.1 -fPIC -w -include cstdint $(INC)

COMMON := $(wildcard $(LIB)/CommonLib/*.cpp) $(wildcard $(LIB)/CommonLib/MotionModels/*.cpp) $(wildcard $(LIB)/CommonLib/x86/*.cpp) \
          $(wildcard $(LIB)/libmd5/*.cpp)
SSE41  := $(wildcard $(LIB)/CommonLib/x86/sse41/*.cpp)
SSE42  := $(wildcard $(LIB)/CommonLib/x86/sse42/*.cpp)
AVX    := $(wildcard $(LIB)/CommonLib/x86/avx/*.cpp)
AVX2   := $(wildcard $(LIB)/CommonLib/x86/avx2/*.cpp)
ENC    := $(wildcard $(LIB)/EncoderLib/*.cpp)
DEC    := $(wildcard $(LIB)/DecoderLib/*.cpp)
UTL    := $(wildcard $(LIB)/Utilities/*.cpp)
EAPP   := $(wildcard $(SRC)/App/EncoderApp/*.cpp)
DAPP   := $(wildcard $(SRC)/App/DecoderApp/*.cpp)

obj = $(patsubst $(SRC)/%.cpp,$(O)/%.o,$(1))
LIBOBJ := $(call obj,$(COMMON) $(SSE41) $(SSE42) $(AVX) $(AVX2) $(UTL))
ENCOBJ := $(call obj,$(ENC))
DECOBJ := $(call obj,$(DEC))

.PHONY: stock gpu dmvr
# the reference's DMVR error-surface functions behind a C entry point (tests/test_oracle_vs_reference_objects.py)
dmvr: $(OUT)/libvtmdmvr.so
$(OUT)/libvtmdmvr.so: $(LIBOBJ) ref_wrap_dmvr.cpp
	$(CXX) $(CXXF) -shared -o $@ ref_wrap_dmvr.cpp $(LIBOBJ) -lm -lpthread
stock: $(OUT)/EncoderApp $(OUT)/DecoderApp

$(O)/Lib/CommonLib/x86/sse41/%.o: $(LIB)/CommonLib/x86/sse41/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXF) -msse4.1 -DUSE_SSE41 -c $< -o $@
$(O)/Lib/CommonLib/x86/sse42/%.o: $(LIB)/CommonLib/x86/sse42/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXF) -msse4.2 -DUSE_SSE42 -c $< -o $@
$(O)/Lib/CommonLib/x86/avx/%.o: $(LIB)/CommonLib/x86/avx/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXF) -mavx -DUSE_AVX -c $< -o $@
$(O)/Lib/CommonLib/x86/avx2/%.o: $(LIB)/CommonLib/x86/avx2/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXF) -mavx2 -DUSE_AVX2 -c $< -o $@
$(O)/%.o: $(SRC)/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXF) -c $< -o $@

$(OUT)/EncoderApp: $(LIBOBJ) $(ENCOBJ) $(DECOBJ) $(call obj,$(EAPP))
	$(CXX) -o $@ $^ -lm -lpthread
$(OUT)/DecoderApp: $(LIBOBJ) $(DECOBJ) $(call obj,$(DAPP))
	$(CXX) -o $@ $^ -lm -lpthread

# ---- GPU-patched apps ---------------------------------------------------------------------------------------------------
# The patch touches .cpp files only (class layouts unchanged), so only these units are rebuilt; the glue unit is ours.
PATCHED := Lib/CommonLib/InterPrediction Lib/CommonLib/MVReprojection Lib/CommonLib/Picture Lib/EncoderLib/InterSearch
P       := $(OUT)/patched
MMROOT  := $(abspath ..)
MMLIBD  := $(MMROOT)/vvc-extension-mm_b200
GPUINC  := -I$(MMROOT)/include -I$(MMLIBD)/host -Ipatches
POBJ    := $(patsubst %,$(P)/%.o,$(PATCHED)) $(P)/MmGpuGlue.o

$(P)/.applied: patches/mmgpu.patch
	rm -rf $(P) && mkdir -p $(P)/Lib/CommonLib $(P)/Lib/EncoderLib
	for f in $(PATCHED); do cp $(SRC)/$$f.cpp $(P)/$$f.cpp; done
	cd $(P) && patch -p1 < $(abspath patches/mmgpu.patch)
	touch $@
$(P)/%.cpp: $(P)/.applied ;
$(P)/MmGpuGlue.o: patches/MmGpuGlue.cpp patches/MmGpuGlue.h $(MMROOT)/include/mmgpu.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXF) $(GPUINC) -c patches/MmGpuGlue.cpp -o $@
$(P)/%.o: $(P)/.applied patches/MmGpuGlue.h
	$(CXX) $(CXXF) $(GPUINC) -c $(P)/$*.cpp -o $@

STOCKOF := $(patsubst %,$(O)/%.o,$(PATCHED))
gpu: $(OUT)/EncoderApp_gpu $(OUT)/DecoderApp_gpu
$(OUT)/EncoderApp_gpu: $(LIBOBJ) $(ENCOBJ) $(DECOBJ) $(call obj,$(EAPP)) $(POBJ)
	$(CXX) -o $@ $(filter-out $(STOCKOF),$(LIBOBJ) $(ENCOBJ) $(DECOBJ) $(call obj,$(EAPP))) $(POBJ) \
	   -L$(MMLIBD) -lmmgpu -Wl,-rpath,'$$ORIGIN/../../../vvc-extension-mm_b200' -lm -lpthread
$(OUT)/DecoderApp_gpu: $(LIBOBJ) $(DECOBJ) $(call obj,$(DAPP)) $(POBJ)
	$(CXX) -o $@ $(filter-out $(STOCKOF) $(O)/Lib/EncoderLib/%,$(LIBOBJ) $(DECOBJ) $(call obj,$(DAPP))) \
	   $(filter-out $(P)/Lib/EncoderLib/%,$(POBJ)) \
	   -L$(MMLIBD) -lmmgpu -Wl,-rpath,'$$ORIGIN/../../../vvc-extension-mm_b200' -lm -lpthread

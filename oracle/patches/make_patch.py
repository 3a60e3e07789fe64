"""Regenerates oracle/patches/mmgpu.patch: copies the four touched reference units to a scratch directory, applies the
edits below (exact-match replacements, each asserted to match once) and writes the unified diff.  The patch, not this
script, is what oracle/vtm.mk applies; the script documents every edit and fails loudly if the reference text moves."""
import difflib
import os
import shutil
import sys

SRC = "/root/reference/source"
FILES = ["Lib/CommonLib/InterPrediction.cpp", "Lib/CommonLib/MVReprojection.cpp", "Lib/CommonLib/Picture.cpp",
         "Lib/EncoderLib/InterSearch.cpp"]
HERE = os.path.dirname(os.path.abspath(__file__))


def sub(text, old, new, count=1):
    assert text.count(old) == count, "expected %d match(es) of %r, found %d" % (count, old[:60], text.count(old))
    return text.replace(old, new)


def edit_picture(t):
    t = sub(t, '#include "CommonLib/InterpolationFilter.h"\n', '#include "CommonLib/InterpolationFilter.h"\n#include "MmGpuGlue.h"\n')
    # the reconstruction changed and its borders were rebuilt: the device copy (if any) is stale
    t = sub(t, "  m_extendedBorder = true;\n}\n", "  m_extendedBorder = true;\n  mmglue::invalidate( this );\n}\n")
    return t


def edit_mvreprojection(t):
    t = sub(t, '#include "MVReprojection.h"\n', '#include "MVReprojection.h"\n#include "MmGpuGlue.h"\n')
    t = sub(t, "  m_initialized = true;\n}\n", "  m_initialized = true;\n  mmglue::init( resolution, sps, epipoleList );\n}\n")
    # MM-MVP candidate conversion through the library (host code of libmmgpu.so, 8f-1)
    t = sub(t, """                                                    const Position &currentBlockPos, const Size &currentBlockSize) const {
  if (motionVectorOrig.hor == 0 && motionVectorOrig.ver == 0) {""", """                                                    const Position &currentBlockPos, const Size &currentBlockSize) const {
  if( mmglue::enabled() && !mmglue::inVerify() && mmglue::mvpOnLibrary() )
  {
    const Mv viaLibrary = mmglue::mvpConvert( position, motionVectorOrig, motionModelIDOrig, motionModelIDDesired, shiftHor, shiftVer, curPOCOrig, refPOCOrig,
                                              curPOCDesired, refPOCDesired, candidateBlockPos, candidateBlockSize, currentBlockPos, currentBlockSize );
    if( mmglue::verifying() )
    {
      mmglue::setInVerify( true );
      const Mv viaReference = motionVectorInDesiredMotionModel( position, motionVectorOrig, motionModelIDOrig, motionModelIDDesired, shiftHor, shiftVer, curPOCOrig,
                                                                refPOCOrig, curPOCDesired, refPOCDesired, candidateBlockPos, candidateBlockSize, currentBlockPos,
                                                                currentBlockSize );
      mmglue::setInVerify( false );
      mmglue::verifyMv( viaLibrary, viaReference );
    }
    return viaLibrary;
  }
  if (motionVectorOrig.hor == 0 && motionVectorOrig.ver == 0) {""")
    return t


def edit_interprediction(t):
    t = sub(t, '#include "debug_tools.h"\n', '#include "debug_tools.h"\n#include "MmGpuGlue.h"\n')
    t = sub(t, """  const Position blockPos = pu.blocks[compID].pos();
  const Size blockSize = pu.blocks[compID].size();
#if INTERPRED_PROFILING
  auto start_mvReprojTime = std::chrono::high_resolution_clock::now();
#endif
  ArrayXXFixedPtrPair cart2DProjMovedFixedSubblocks = m_mvReprojection->reprojectMotionVectorSubblocks(
    blockPos, blockSize, mv, motionModel, compID, chFmt, pu.cs->slice->getPOC(), refPic->getPOC());
""", """  const Position blockPos = pu.blocks[compID].pos();
  const Size blockSize = pu.blocks[compID].size();
  if( mmglue::enabled() && !mmglue::inVerify() && !srcPadBuf && !bilinearMC )
  {
    // GPU path: mapping + interpolation of the whole block in one call (libmmgpu.so, mmgpu_mc)
    mmglue::predInterBlkMM( compID, chFmt, blockPos, blockSize, refPic, wrapRef, mv, motionModel, bi, clpRng, pu.cs->slice->getPOC(),
                            dstPic.bufs[compID] );
    if( mmglue::verifying() )
    {
      // MMGPU_VERIFY=1: the reference's own code on the same call, compared sample by sample
      mmglue::verifySave( dstPic.bufs[compID].buf, (int) dstPic.bufs[compID].stride, (int) blockSize.width, (int) blockSize.height );
      mmglue::setInVerify( true );
      xPredInterBlkMM( compID, pu, refPic, _mv, dstPic, motionModel, bi, clpRng, bdofApplied, isIBC, scalingRation, bilinearMC, srcPadBuf, srcPadStride );
      mmglue::setInVerify( false );
      mmglue::verifyCompare( dstPic.bufs[compID].buf, (int) dstPic.bufs[compID].stride, (int) blockSize.width, (int) blockSize.height, "xPredInterBlkMM" );
    }
    return;
  }
#if INTERPRED_PROFILING
  auto start_mvReprojTime = std::chrono::high_resolution_clock::now();
#endif
  ArrayXXFixedPtrPair cart2DProjMovedFixedSubblocks = m_mvReprojection->reprojectMotionVectorSubblocks(
    blockPos, blockSize, mv, motionModel, compID, chFmt, pu.cs->slice->getPOC(), refPic->getPOC());
""")
    # classic translational MC (plain case) on the device as well
    t = sub(t, """    PelBuf & dstBuf = dstPic.bufs[compID];
    unsigned width  = dstBuf.width;
    unsigned height = dstBuf.height;

    CPelBuf refBuf;
    {
      Position offset = pu.blocks[compID].pos().offset(mv.getHor() >> shiftHor, mv.getVer() >> shiftVer);""", """    PelBuf & dstBuf = dstPic.bufs[compID];
    unsigned width  = dstBuf.width;
    unsigned height = dstBuf.height;
    if( mmglue::enabled() && !mmglue::inVerify() && mmglue::classicOnGpu() && !isIBC && !bioApplied && !bilinearMC && !dmvrWidth && !srcPadBuf && !useAltHpelIf
        && width >= 4 && height >= 4 && ( width & ( width - 1 ) ) == 0 && ( height & ( height - 1 ) ) == 0
        && width == pu.blocks[compID].width && height == pu.blocks[compID].height )
    {
      mmglue::predInterBlkClassic( compID, chFmt, pu.blocks[compID].pos(), int( width ), int( height ), refPic, wrapRef, mv, bi, clpRng, dstBuf );
      if( mmglue::verifying() )
      {
        mmglue::verifySave( dstBuf.buf, (int) dstBuf.stride, (int) width, (int) height );
        mmglue::setInVerify( true );
        xPredInterBlk( compID, pu, refPic, _mv, dstPic, bi, clpRng, bioApplied, isIBC, scalingRatio, dmvrWidth, dmvrHeight, bilinearMC, srcPadBuf, srcPadStride );
        mmglue::setInVerify( false );
        mmglue::verifyCompare( dstBuf.buf, (int) dstBuf.stride, (int) width, (int) height, "xPredInterBlk" );
      }
      return;
    }
    if( mmglue::enabled() && !mmglue::inVerify() && mmglue::classicOnGpu() && !isIBC && bioApplied && compID == COMPONENT_Y && bi && !bilinearMC && !dmvrWidth && !srcPadBuf
        && !useAltHpelIf && width >= 8 && height >= 8 && ( width & ( width - 1 ) ) == 0 && ( height & ( height - 1 ) ) == 0
        && width == pu.blocks[compID].width && height == pu.blocks[compID].height )
    {
      // prediction + BDOF border fetch into the work buffer, as the code below does
      mmglue::predInterBlkClassicBdof( pu.blocks[compID].pos(), int( width ), int( height ), refPic, wrapRef, mv, clpRng,
                                       m_filteredBlockTmp[2 + m_iRefListIdx][compID] );
      if( mmglue::verifying() )
      {
        // the ring box: rows 1 .. h+2, columns 1 .. w+2 of the (w + 4)-stride work buffer
        Pel* box = m_filteredBlockTmp[2 + m_iRefListIdx][compID] + ( width + 4 ) + 1;
        mmglue::verifySave( box, (int) width + 4, (int) width + 2, (int) height + 2 );
        mmglue::setInVerify( true );
        xPredInterBlk( compID, pu, refPic, _mv, dstPic, bi, clpRng, bioApplied, isIBC, scalingRatio, dmvrWidth, dmvrHeight, bilinearMC, srcPadBuf, srcPadStride );
        mmglue::setInVerify( false );
        mmglue::verifyCompare( box, (int) width + 4, (int) width + 2, (int) height + 2, "xPredInterBlk (BDOF)" );
      }
      return;
    }

    CPelBuf refBuf;
    {
      Position offset = pu.blocks[compID].pos().offset(mv.getHor() >> shiftHor, mv.getVer() >> shiftVer);""")
    # luma with the alternative half-pel filter (pu.cu->imv == IMV_HPEL): own entry point; chroma ignores the flag
    t = sub(t, """    if( mmglue::enabled() && !mmglue::inVerify() && mmglue::classicOnGpu() && !isIBC && !bioApplied && !bilinearMC && !dmvrWidth && !srcPadBuf && !useAltHpelIf
        && width >= 4""", """    if( mmglue::enabled() && !mmglue::inVerify() && mmglue::classicOnGpu() && !isIBC && !bioApplied && !bilinearMC && !dmvrWidth && !srcPadBuf && useAltHpelIf
        && compID == COMPONENT_Y && width >= 4 && height >= 4 && ( width % 4 ) == 0 && ( height % 4 ) == 0
        && width == pu.blocks[compID].width && height == pu.blocks[compID].height )
    {
      mmglue::predInterBlkClassicAltHpel( pu.blocks[compID].pos(), int( width ), int( height ), refPic, wrapRef, mv, bi, clpRng, dstBuf );
      if( mmglue::verifying() )
      {
        mmglue::verifySave( dstBuf.buf, (int) dstBuf.stride, (int) width, (int) height );
        mmglue::setInVerify( true );
        xPredInterBlk( compID, pu, refPic, _mv, dstPic, bi, clpRng, bioApplied, isIBC, scalingRatio, dmvrWidth, dmvrHeight, bilinearMC, srcPadBuf, srcPadStride );
        mmglue::setInVerify( false );
        mmglue::verifyCompare( dstBuf.buf, (int) dstBuf.stride, (int) width, (int) height, "xPredInterBlk (alternative half-pel filter)" );
      }
      return;
    }
    if( mmglue::enabled() && !mmglue::inVerify() && mmglue::classicOnGpu() && !isIBC && !bioApplied && !bilinearMC && !dmvrWidth && !srcPadBuf
        && ( !useAltHpelIf || compID != COMPONENT_Y ) && width >= 4""")
    # ---- classic DMVR (xProcessDMVR): integer search of all sub-PUs in one batch, final MC of a sub-PU in one batch
    t = sub(t, """    if (*(pSADsArray + sadOffset) == MAX_UINT64)
    {
      const uint64_t cost = xDMVRCost(bd, pRefL0, refStrideL0, pRefL1, refStrideL1, width, height);""", """    if (*(pSADsArray + sadOffset) == MAX_UINT64)
    {
      const uint64_t cost = mmglue::dmvrActive() ? mmglue::dmvrCost( nIdx ) : xDMVRCost(bd, pRefL0, refStrideL0, pRefL1, refStrideL1, width, height);""")
    t = sub(t, """    srcPred0 = srcPred0.subBuf(UnitAreaRelative(pu, subPu));
    srcPred1 = srcPred1.subBuf(UnitAreaRelative(pu, subPu));

    int yStart = 0;
    for (int y = puPos.y; y < (puPos.y + pu.lumaSize().height); y = y + dy, yStart = yStart + dy)
    {
      for (int x = puPos.x, xStart = 0; x < (puPos.x + pu.lumaSize().width); x = x + dx, xStart = xStart + dx)
      {
        PredictionUnit subPu = pu;
        subPu.UnitArea::operator=(UnitArea(pu.chromaFormat, Area(x, y, dx, dy)));
        xPrefetch(subPu, m_cYuvRefBuffDMVRL0, REF_PIC_LIST_0, 1);
        xPrefetch(subPu, m_cYuvRefBuffDMVRL1, REF_PIC_LIST_1, 1);

        xinitMC(subPu, clpRngs);

        uint64_t minCost = MAX_UINT64;""", """    srcPred0 = srcPred0.subBuf(UnitAreaRelative(pu, subPu));
    srcPred1 = srcPred1.subBuf(UnitAreaRelative(pu, subPu));

    const bool gpuDmvr = mmglue::enabled() && mmglue::classicOnGpu() && mmglue::dmvrOnGpu() && ( dx == 8 || dx == 16 ) && ( dy == 8 || dy == 16 );
    if( gpuDmvr )
    {
      // the bilinear integer search of every sub-PU in ONE device batch: per list the window origin of xPrefetch and the
      // fraction xinitMC's xPredInterBlk filters with (same statements as there)
      mmglue::dmvrBegin( dx, dy );
      for( int y = puPos.y; y < ( puPos.y + pu.lumaSize().height ); y = y + dy )
      {
        for( int x = puPos.x; x < ( puPos.x + pu.lumaSize().width ); x = x + dx )
        {
          PredictionUnit sp = pu;
          sp.UnitArea::operator=( UnitArea( pu.chromaFormat, Area( x, y, dx, dy ) ) );
          mmglue::DmvrList l[2];
          for( int k = 0; k < NUM_REF_PIC_LIST_01; k++ )
          {
            const RefPicList refId  = RefPicList( k );
            const Picture*   refPic = sp.cu->slice->getRefPic( refId, sp.refIdx[refId] )->unscaledPic;
            Mv cMv( sp.mv[refId].getHor(), sp.mv[refId].getVer() );
            cMv += Mv( -( ( ( NTAPS_LUMA >> 1 ) - 1 ) << MV_FRACTIONAL_BITS_INTERNAL ), -( ( ( NTAPS_LUMA >> 1 ) - 1 ) << MV_FRACTIONAL_BITS_INTERNAL ) );
            bool wrapRef = false;
            if( refPic->isWrapAroundEnabled( sp.cs->pps ) )
            {
              wrapRef = wrapClipMv( cMv, sp.blocks[0].pos(), sp.blocks[0].size(), sp.cs->sps, sp.cs->pps, CLASSIC );
            }
            else
            {
              clipMv( cMv, sp.lumaPos(), sp.lumaSize(), *sp.cs->sps, *sp.cs->pps, CLASSIC );
            }
            Mv mergeMv( sp.mv[refId] );
            if( !sp.cs->pps->getWrapAroundEnabledFlag() )
            {
              clipMv( mergeMv, sp.lumaPos(), sp.lumaSize(), *sp.cs->sps, *sp.cs->pps, CLASSIC );
            }
            if( refPic->isWrapAroundEnabled( sp.cs->pps ) )
            {
              wrapClipMv( mergeMv, sp.blocks[0].pos(), sp.blocks[0].size(), sp.cs->sps, sp.cs->pps, CLASSIC );
            }
            l[k] = mmglue::DmvrList{ refPic, wrapRef, x + ( cMv.getHor() >> MV_FRACTIONAL_BITS_INTERNAL ), y + ( cMv.getVer() >> MV_FRACTIONAL_BITS_INTERNAL ),
                                     mergeMv.hor & 15, mergeMv.ver & 15 };
          }
          mmglue::dmvrAdd( l[0], l[1] );
        }
      }
      mmglue::dmvrRun();
    }

    int yStart = 0;
    for (int y = puPos.y; y < (puPos.y + pu.lumaSize().height); y = y + dy, yStart = yStart + dy)
    {
      for (int x = puPos.x, xStart = 0; x < (puPos.x + pu.lumaSize().width); x = x + dx, xStart = xStart + dx)
      {
        PredictionUnit subPu = pu;
        subPu.UnitArea::operator=(UnitArea(pu.chromaFormat, Area(x, y, dx, dy)));
        xPrefetch(subPu, m_cYuvRefBuffDMVRL0, REF_PIC_LIST_0, 1);
        xPrefetch(subPu, m_cYuvRefBuffDMVRL1, REF_PIC_LIST_1, 1);

        if( gpuDmvr )
        {
          mmglue::dmvrSelect( num );
          if( mmglue::verifying() )
          {
            // the reference's own bilinear predictions and its 25 costs next to the device's
            xinitMC( subPu, clpRngs );
            for( int nIdx = 0; nIdx < 25; nIdx++ )
            {
              Pel* a0 = biLinearPredL0 + m_pSearchOffset[nIdx].hor + ( m_pSearchOffset[nIdx].ver * m_biLinearBufStride );
              Pel* a1 = biLinearPredL1 - m_pSearchOffset[nIdx].hor - ( m_pSearchOffset[nIdx].ver * m_biLinearBufStride );
              mmglue::dmvrVerifyCost( nIdx, xDMVRCost( bd, a0, m_biLinearBufStride, a1, m_biLinearBufStride, dx, dy ) );
            }
          }
        }
        else
        {
          xinitMC(subPu, clpRngs);
        }

        uint64_t minCost = MAX_UINT64;""")
    t = sub(t, """            minCost = xDMVRCost(clpRngs.comp[COMPONENT_Y].bd, addrL0, m_biLinearBufStride, addrL1, m_biLinearBufStride, dx, dy);
            minCost -= (minCost >>2);
            if (minCost < (dx * dy))
            {
              notZeroCost = false;
              break;
            }
            pSADsArray[0] = minCost;
          }
          if (!minCost)
          {
            notZeroCost = false;
            break;
          }

          xBIPMVRefine(bd, addrL0, addrL1, minCost, deltaMV, pSADsArray, dx, dy);""", """            minCost = gpuDmvr ? mmglue::dmvrCost( 12 ) : xDMVRCost(clpRngs.comp[COMPONENT_Y].bd, addrL0, m_biLinearBufStride, addrL1, m_biLinearBufStride, dx, dy);
            minCost -= (minCost >>2);
            if (minCost < (dx * dy))
            {
              notZeroCost = false;
              break;
            }
            pSADsArray[0] = minCost;
          }
          if (!minCost)
          {
            notZeroCost = false;
            break;
          }

          xBIPMVRefine(bd, addrL0, addrL1, minCost, deltaMV, pSADsArray, dx, dy);""")
    t = sub(t, """        xWeightedAverage(subPu, srcPred0, srcPred1, subPredBuf, subPu.cu->slice->getSPS()->getBitDepths(), subPu.cu->slice->clpRngs(), bioAppliedType[num]);
        num++;
      }
    }
  }
  JVET_J0090_SET_CACHE_ENABLE(true);
}""", """        xWeightedAverage(subPu, srcPred0, srcPred1, subPredBuf, subPu.cu->slice->getSPS()->getBitDepths(), subPu.cu->slice->clpRngs(), bioAppliedType[num]);
        num++;
      }
    }
    mmglue::dmvrEnd();
  }
  JVET_J0090_SET_CACHE_ENABLE(true);
}""")
    # xFinalPaddedMCForDMVR: the components that read the padded window go to the device in one batch per sub-PU
    t = sub(t, """        offset = (DMVR_NUM_ITERATION + leftPixelExtra) * (pcPadTemp.bufs[compID].stride + 1);
        offset += (deltaIntMvY)* pcPadTemp.bufs[compID].stride;
        offset += (deltaIntMvX);
        srcBufPelPtr = (srcBuf.buf + offset);
      }
""", """        offset = (DMVR_NUM_ITERATION + leftPixelExtra) * (pcPadTemp.bufs[compID].stride + 1);
        offset += (deltaIntMvY)* pcPadTemp.bufs[compID].stride;
        offset += (deltaIntMvX);
        srcBufPelPtr = (srcBuf.buf + offset);
        const int cw = pu.blocks[compID].width, ch = pu.blocks[compID].height;
        if( mmglue::enabled() && mmglue::classicOnGpu() && mmglue::dmvrOnGpu() && cw <= 16 && ch <= 16
            && cw == (int) pcYUVTemp.bufs[compID].width && ch == (int) pcYUVTemp.bufs[compID].height )
        {
          // window origin as xPrefetch computed it from the merge MV (InterPrediction.cpp:1906-1931) ...
          const int filtersize = ( compID == COMPONENT_Y ) ? NTAPS_LUMA : NTAPS_CHROMA;
          Mv wMv( startMv.getHor(), startMv.getVer() );
          wMv += Mv( -( ( ( filtersize >> 1 ) - 1 ) << mvshiftTempHor ), -( ( ( filtersize >> 1 ) - 1 ) << mvshiftTempVer ) );
          bool wrapRef = false;
          if( refPic->isWrapAroundEnabled( pu.cs->pps ) )
          {
            wrapRef = wrapClipMv( wMv, pu.blocks[0].pos(), pu.blocks[0].size(), pu.cs->sps, pu.cs->pps, CLASSIC );
          }
          else
          {
            clipMv( wMv, pu.lumaPos(), pu.lumaSize(), *pu.cs->sps, *pu.cs->pps, CLASSIC );
          }
          // ... and the fraction xPredInterBlk filters with (:942-971)
          Mv fMv( cMvClipped );
          if( refPic->isWrapAroundEnabled( pu.cs->pps ) )
          {
            wrapClipMv( fMv, pu.blocks[0].pos(), pu.blocks[0].size(), pu.cs->sps, pu.cs->pps, CLASSIC );
          }
          const int csx = ::getComponentScaleX( (ComponentID) compID, pu.chromaFormat ), csy = ::getComponentScaleY( (ComponentID) compID, pu.chromaFormat );
          const int xFrac = compID == COMPONENT_Y ? ( fMv.hor & 15 ) : ( fMv.hor * ( 1 << ( 1 - csx ) ) & 31 );
          const int yFrac = compID == COMPONENT_Y ? ( fMv.ver & 15 ) : ( fMv.ver * ( 1 << ( 1 - csy ) ) & 31 );
          const bool bio = bioApplied && compID == COMPONENT_Y;
          mmglue::DmvrFinal job = { (ComponentID) compID, refPic, wrapRef,
                                    pu.blocks[compID].x + ( wMv.getHor() >> mvshiftTempHor ), pu.blocks[compID].y + ( wMv.getVer() >> mvshiftTempVer ), cw, ch,
                                    blockMoved ? ( DMVR_NUM_ITERATION >> csy ) : 0, deltaIntMvX, deltaIntMvY, xFrac, yFrac, pu.cu->imv == IMV_HPEL, bio,
                                    bio ? m_filteredBlockTmp[2 + m_iRefListIdx][compID] : pcYUVTemp.bufs[compID].buf, (int) pcYUVTemp.bufs[compID].stride };
          mmglue::dmvrFinalAdd( job );
          if( !mmglue::verifying() )
          {
            continue;
          }
        }
      }
""")
    t = sub(t, """    pcYUVTemp = pcYuvSrc1;
    pcPadTemp = pcPad1;
  }
}

uint64_t InterPrediction::xDMVRCost(""", """    pcYUVTemp = pcYuvSrc1;
    pcPadTemp = pcPad1;
  }
  mmglue::dmvrFinalRun();
}

uint64_t InterPrediction::xDMVRCost(""")
    # debugging aid (MMGPU_DUMP): record the CPU path's result
    t = sub(t, """      }
    }
  }
  JVET_J0090_SET_CACHE_ENABLE(
    (srcPadStride == 0)
    && (bioApplied
        == false));   // Enabled only in non-DMVR-non-BDOF process, In DMVR process, srcPadStride is always non-zero
#if INTERPRED_PROFILING
  auto end_interpolTime = std::chrono::high_resolution_clock::now();
  dbg_interpolTime += std::chrono::duration<double>(end_interpolTime - start_interpolTime).count();
  auto start_bdofPadTime""", """      }
    }
  }
  if( mmglue::dumping() && !srcPadBuf && !bilinearMC )
  {
    mmglue::dumpMc( compID, chFmt, blockPos, blockSize, refPic, wrapRef, mv, motionModel, bi, clpRng, pu.cs->slice->getPOC(), dstBuf );
  }
  JVET_J0090_SET_CACHE_ENABLE(
    (srcPadStride == 0)
    && (bioApplied
        == false));   // Enabled only in non-DMVR-non-BDOF process, In DMVR process, srcPadStride is always non-zero
#if INTERPRED_PROFILING
  auto end_interpolTime = std::chrono::high_resolution_clock::now();
  dbg_interpolTime += std::chrono::duration<double>(end_interpolTime - start_interpolTime).count();
  auto start_bdofPadTime""")
    return t


MEBLK = "mmglue::MeBlock{ %s.blkPos, %s.blkSize, %s.pcRefBuf->buf, %s.motionModel, %s.curPOC, %s.refPOC, m_lumaClpRng }"


def meblk(s):
    return MEBLK % ((s,) * 6)


def edit_intersearch(t):
    t = sub(t, '#include "EncLib.h"\n', '#include "EncLib.h"\n#include "MmGpuGlue.h"\n')
    # --- xTZSearchHelp: during a dry run of an enumerator only the visited point is recorded
    t = sub(t, """  Distortion  uiSad = 0;

//  CHECK(!( !( rcStruct.searchRange.left > iSearchX""", """  Distortion  uiSad = 0;
  if( mmglue::collecting() )
  {
    mmglue::collect( iSearchX, iSearchY );
    return;
  }

//  CHECK(!( !( rcStruct.searchRange.left > iSearchX""")
    # --- xTZ2PointSearch: both points ahead
    t = sub(t, """  const int iY2 = rcStruct.iBestY + yOffset[1][rcStruct.ucPointNr];

""", """  const int iY2 = rcStruct.iBestY + yOffset[1][rcStruct.ucPointNr];
  if( mmglue::enabled() && rcStruct.motionModel != CLASSIC && !mmglue::collecting() )
  {
    const Mv ahead[2] = { Mv( iX1, iY1 ), Mv( iX2, iY2 ) };
    const bool in1 = iX1 >= sr.left && iX1 <= sr.right && iY1 >= sr.top && iY1 <= sr.bottom;
    const bool in2 = iX2 >= sr.left && iX2 <= sr.right && iY2 >= sr.top && iY2 <= sr.bottom;
    if( in1 || in2 )
    {
      mmglue::mePrefetch( %s, m_cDistParam, ahead + ( in1 ? 0 : 1 ), ( in1 && in2 ) ? 2 : 1, MV_PRECISION_INT );
    }
  }

""" % meblk("rcStruct"))
    # --- xMotionEstimation: the reference picture must be resident before the search starts
    t = sub(t, """    buf = refPic->getRecoBuf(COMPONENT_Y, wrap);
  }

  IntTZSearchStruct cStruct;""", """    buf = refPic->getRecoBuf(COMPONENT_Y, wrap);
    if( mmglue::enabled() )
    {
      mmglue::ensureRef( refPic, wrap );
    }
    else if( mmglue::dumping() )
    {
      mmglue::dumpRef( refPic, wrap );
    }
  }

  IntTZSearchStruct cStruct;""")
    # --- xPatternSearch: the whole grid (bi-prediction: 81 points) in one batch
    t = sub(t, """  const SearchRange& sr = cStruct.searchRange;

  for ( int y = sr.top; y <= sr.bottom; y++ )
  {
    for ( int x = sr.left; x <= sr.right; x++ )
    {
      //  find min. distortion position
      xApplyMvVA(cStruct, Mv(x, y), MV_PRECISION_INT);
""", """  const SearchRange& sr = cStruct.searchRange;

  if( mmglue::enabled() && cStruct.motionModel != CLASSIC )
  {
    mmglue::beginCollect();
    for ( int y = sr.top; y <= sr.bottom; y++ )
    {
      for ( int x = sr.left; x <= sr.right; x++ )
      {
        mmglue::collect( x, y );
      }
    }
    mmglue::endCollect( %s, m_cDistParam );
  }
  for ( int y = sr.top; y <= sr.bottom; y++ )
  {
    for ( int x = sr.left; x <= sr.right; x++ )
    {
      //  find min. distortion position
      xApplyMvVA(cStruct, Mv(x, y), MV_PRECISION_INT);
""" % meblk("cStruct"))
    # --- xTZSearch: first search -- dry run of all diamond rounds around the start point, one batch
    t = sub(t, """  // first search around best position up to now.
  // The following works as a "subsampled/log" window search around the best candidate
  for (iDist = 1; iDist <= searchRange; iDist *= 2)
  {
""", """  // first search around best position up to now.
  // The following works as a "subsampled/log" window search around the best candidate
  if( mmglue::enabled() && cStruct.motionModel != CLASSIC )
  {
    // look-ahead: dry run of the reference's own enumerators (xTZSearchHelp only records the points), one GPU batch
    IntTZSearchStruct dry = cStruct;
    mmglue::beginCollect();
    for (iDist = 1; iDist <= searchRange; iDist *= 2)
    {
      if ( bFirstSearchDiamond == 1 )
      {
        xTZ8PointDiamondSearch ( dry, iStartX, iStartY, iDist, bFirstCornersForDiamondDist1 );
      }
      else
      {
        xTZ8PointSquareSearch  ( dry, iStartX, iStartY, iDist );
      }
    }
    mmglue::endCollect( %s, m_cDistParam );
  }
  for (iDist = 1; iDist <= searchRange; iDist *= 2)
  {
""" % meblk("cStruct"))
    # --- xTZSearch: raster searches
    t = sub(t, """    cStruct.uiBestDistance = iWindowSize;
    for ( iStartY = localsr.top; iStartY <= localsr.bottom; iStartY += iWindowSize )
    {
""", """    cStruct.uiBestDistance = iWindowSize;
    if( mmglue::enabled() && cStruct.motionModel != CLASSIC )
    {
      mmglue::beginCollect();
      for ( iStartY = localsr.top; iStartY <= localsr.bottom; iStartY += iWindowSize )
      {
        for ( iStartX = localsr.left; iStartX <= localsr.right; iStartX += iWindowSize )
        {
          mmglue::collect( iStartX, iStartY );
        }
      }
      mmglue::endCollect( %s, m_cDistParam );
    }
    for ( iStartY = localsr.top; iStartY <= localsr.bottom; iStartY += iWindowSize )
    {
""" % meblk("cStruct"))
    t = sub(t, """      cStruct.uiBestDistance = iRaster;
      for ( iStartY = sr.top; iStartY <= sr.bottom; iStartY += iRaster )
      {
""", """      cStruct.uiBestDistance = iRaster;
      if( mmglue::enabled() && cStruct.motionModel != CLASSIC )
      {
        mmglue::beginCollect();
        for ( iStartY = sr.top; iStartY <= sr.bottom; iStartY += iRaster )
        {
          for ( iStartX = sr.left; iStartX <= sr.right; iStartX += iRaster )
          {
            mmglue::collect( iStartX, iStartY );
          }
        }
        mmglue::endCollect( %s, m_cDistParam );
      }
      for ( iStartY = sr.top; iStartY <= sr.bottom; iStartY += iRaster )
      {
""" % meblk("cStruct"))
    # --- xTZSearch: star refinement -- all rounds around the new start point
    t = sub(t, """      cStruct.uiBestDistance = 0;
      cStruct.ucPointNr = 0;
      for (iDist = 1; iDist < searchRange + 1; iDist *= 2)
      {
        if ( bStarRefinementDiamond == 1 )
        {
          xTZ8PointDiamondSearch ( cStruct, iStartX, iStartY, iDist, bStarRefinementCornersForDiamondDist1 );
""", """      cStruct.uiBestDistance = 0;
      cStruct.ucPointNr = 0;
      if( mmglue::enabled() && cStruct.motionModel != CLASSIC )
      {
        IntTZSearchStruct dry = cStruct;
        mmglue::beginCollect();
        for (iDist = 1; iDist < searchRange + 1; iDist *= 2)
        {
          if ( bStarRefinementDiamond == 1 )
          {
            xTZ8PointDiamondSearch ( dry, iStartX, iStartY, iDist, bStarRefinementCornersForDiamondDist1 );
          }
          else
          {
            xTZ8PointSquareSearch  ( dry, iStartX, iStartY, iDist );
          }
        }
        mmglue::endCollect( %s, m_cDistParam );
      }
      for (iDist = 1; iDist < searchRange + 1; iDist *= 2)
      {
        if ( bStarRefinementDiamond == 1 )
        {
          xTZ8PointDiamondSearch ( cStruct, iStartX, iStartY, iDist, bStarRefinementCornersForDiamondDist1 );
""" % meblk("cStruct"), count=1)
    # --- fractional refinement: distortion set-up + the 9 points of a refinement step in one batch
    t = sub(t, """    m_pcRdCost->setDistParam(m_cDistParam, *cStruct.pcPatternKey, m_tmpMMStorage.buf, m_tmpMMStorage.stride, m_lumaClpRng.bd, COMPONENT_Y, 0, 1, m_pcEncCfg->getUseHADME() && !pu.cs->slice->getDisableSATDForRD());
    xMVReprojectionInterpolation(cStruct.blkPos, cStruct.blkSize, *cStruct.pcRefBuf, rcMvBaseI, MV_PRECISION_INT,
                                 m_tmpMMStorage, cStruct.motionModel, m_lumaClpRng, cStruct.curPOC, cStruct.refPOC);
""", """    m_pcRdCost->setDistParam(m_cDistParam, *cStruct.pcPatternKey, m_tmpMMStorage.buf, m_tmpMMStorage.stride, m_lumaClpRng.bd, COMPONENT_Y, 0, 1, m_pcEncCfg->getUseHADME() && !pu.cs->slice->getDisableSATDForRD());
    if( mmglue::dumping() )
    {
      mmglue::meDumpSetup( m_cDistParam, m_pcEncCfg->getUseHADME() && !pu.cs->slice->getDisableSATDForRD() );
    }
    if( mmglue::enabled() )
    {
      mmglue::meSetup( m_cDistParam, m_pcEncCfg->getUseHADME() && !pu.cs->slice->getDisableSATDForRD() );
      mmglue::meEval( %s, m_cDistParam, rcMvBaseI, MV_PRECISION_INT );
      if( mmglue::verifying() )
      {
        xMVReprojectionInterpolation(cStruct.blkPos, cStruct.blkSize, *cStruct.pcRefBuf, rcMvBaseI, MV_PRECISION_INT,
                                     m_tmpMMStorage, cStruct.motionModel, m_lumaClpRng, cStruct.curPOC, cStruct.refPOC);
        mmglue::meVerify( m_cDistParam );
      }
    }
    else
    xMVReprojectionInterpolation(cStruct.blkPos, cStruct.blkSize, *cStruct.pcRefBuf, rcMvBaseI, MV_PRECISION_INT,
                                 m_tmpMMStorage, cStruct.motionModel, m_lumaClpRng, cStruct.curPOC, cStruct.refPOC);
""" % meblk("cStruct"))
    t = sub(t, """  m_pcRdCost->setDistParam( m_cDistParam, *cStruct.pcPatternKey, m_tmpMMStorage.buf, m_tmpMMStorage.stride, m_lumaClpRng.bd, COMPONENT_Y, 0, 1, useHadamard );

  // Half-pel refinement
""", """  m_pcRdCost->setDistParam( m_cDistParam, *cStruct.pcPatternKey, m_tmpMMStorage.buf, m_tmpMMStorage.stride, m_lumaClpRng.bd, COMPONENT_Y, 0, 1, useHadamard );
  if( mmglue::enabled() )
  {
    mmglue::meSetup( m_cDistParam, useHadamard );
  }
  else if( mmglue::dumping() )
  {
    mmglue::meDumpSetup( m_cDistParam, useHadamard );
  }

  // Half-pel refinement
""")
    t = sub(t, """  const Mv* pcMvRefine = (refinementPrecision == MV_PRECISION_HALF ? s_acMvRefineH : s_acMvRefineQ);
  for (uint32_t i = 0; i < 9; ++i) {
    cMvTest = cMvBase + pcMvRefine[i];
    xMVReprojectionInterpolation(cStruct.blkPos, cStruct.blkSize, *cStruct.pcRefBuf, cMvTest, refinementPrecision,
                                 m_tmpMMStorage, cStruct.motionModel, m_lumaClpRng, cStruct.curPOC, cStruct.refPOC);
""", """  const Mv* pcMvRefine = (refinementPrecision == MV_PRECISION_HALF ? s_acMvRefineH : s_acMvRefineQ);
  if( mmglue::enabled() )
  {
    Mv ahead[9];
    for (uint32_t i = 0; i < 9; ++i) {
      ahead[i] = cMvBase + pcMvRefine[i];
    }
    mmglue::mePrefetch( %s, m_cDistParam, ahead, 9, refinementPrecision );
  }
  for (uint32_t i = 0; i < 9; ++i) {
    cMvTest = cMvBase + pcMvRefine[i];
    if( mmglue::enabled() )
    {
      mmglue::meEval( %s, m_cDistParam, cMvTest, refinementPrecision );
      if( mmglue::verifying() )
      {
        xMVReprojectionInterpolation(cStruct.blkPos, cStruct.blkSize, *cStruct.pcRefBuf, cMvTest, refinementPrecision,
                                     m_tmpMMStorage, cStruct.motionModel, m_lumaClpRng, cStruct.curPOC, cStruct.refPOC);
        mmglue::meVerify( m_cDistParam );
      }
    }
    else
    xMVReprojectionInterpolation(cStruct.blkPos, cStruct.blkSize, *cStruct.pcRefBuf, cMvTest, refinementPrecision,
                                 m_tmpMMStorage, cStruct.motionModel, m_lumaClpRng, cStruct.curPOC, cStruct.refPOC);
""" % (meblk("cStruct"), meblk("cStruct")))
    # --- prepareRdCostDistParamVA / xApplyMvVA: the MM branches
    t = sub(t, """    m_pcRdCost->setDistParam(m_cDistParam, *rcStruct.pcPatternKey, m_tmpMMStorage.buf, m_tmpMMStorage.stride, m_lumaClpRng.bd, COMPONENT_Y, subShiftMode, step, useHadamard);
  }
}
""", """    m_pcRdCost->setDistParam(m_cDistParam, *rcStruct.pcPatternKey, m_tmpMMStorage.buf, m_tmpMMStorage.stride, m_lumaClpRng.bd, COMPONENT_Y, subShiftMode, step, useHadamard);
    if( mmglue::enabled() )
    {
      CHECK( step != 1, "mmglue: distortion step != 1" );
      mmglue::meSetup( m_cDistParam, useHadamard );
    }
    else if( mmglue::dumping() )
    {
      mmglue::meDumpSetup( m_cDistParam, useHadamard );
    }
  }
}
""")
    t = sub(t, """  else
  {
    xMVReprojectionInterpolation(rcStruct.blkPos, rcStruct.blkSize, *rcStruct.pcRefBuf, rMv, mvPrec, m_tmpMMStorage, rcStruct.motionModel, m_lumaClpRng, rcStruct.curPOC, rcStruct.refPOC);
  }
}
""", """  else if( mmglue::enabled() )
  {
    mmglue::meEval( %s, m_cDistParam, rMv, mvPrec );
    if( mmglue::verifying() )
    {
      xMVReprojectionInterpolation(rcStruct.blkPos, rcStruct.blkSize, *rcStruct.pcRefBuf, rMv, mvPrec, m_tmpMMStorage, rcStruct.motionModel, m_lumaClpRng, rcStruct.curPOC, rcStruct.refPOC);
      mmglue::meVerify( m_cDistParam );
    }
  }
  else
  {
    xMVReprojectionInterpolation(rcStruct.blkPos, rcStruct.blkSize, *rcStruct.pcRefBuf, rMv, mvPrec, m_tmpMMStorage, rcStruct.motionModel, m_lumaClpRng, rcStruct.curPOC, rcStruct.refPOC);
  }
}
""" % meblk("rcStruct"))
    # --- classic fractional refinement: the 9 + 9 candidates of xPatternRefinement as classic MC + distortion on the device
    t = sub(t, """  //  Half-pel refinement
  m_pcRdCost->setCostScale(1);
  xExtDIFUpSamplingH(&cPatternRoi, cStruct.useAltHpelIf);
""", """  //  Half-pel refinement
  m_pcRdCost->setCostScale(1);
  xExtDIFUpSamplingH(&cPatternRoi, cStruct.useAltHpelIf);
  if( ( mmglue::enabled() && mmglue::classicOnGpu() ) || mmglue::dumping() )
  {
    // the reference picture the search runs on: cStruct.pcRefBuf is the block-sized view for CLASSIC (:4993), its plane origin
    // is blkPos samples before it
    const Picture* refPicC = pu.cu->slice->getRefPic( m_currRefPicList, m_currRefPicIndex );
    const bool wrapC = refPicC->isWrapAroundEnabled( pu.cs->pps );
    const CPelBuf planeC = refPicC->getRecoBuf( COMPONENT_Y, wrapC );
    if( mmglue::enabled() ) mmglue::ensureRef( refPicC, wrapC ); else mmglue::dumpRef( refPicC, wrapC );
    mmglue::classicFracBegin( mmglue::MeBlock{ cStruct.blkPos, cStruct.blkSize, planeC.buf, CLASSIC, cStruct.curPOC, cStruct.refPOC, m_lumaClpRng },
                              rcMvBaseI, !cStruct.useAltHpelIf );
  }
""")
    t = sub(t, """    rcMvRefinedQ += cMvQter;
  }
}
""", """    rcMvRefinedQ += cMvQter;
  }
  mmglue::classicFracEnd();
}
""")
    t = sub(t, """  m_pcRdCost->setDistParam( m_cDistParam, *pcPatternKey, m_filteredBlock[0][0][0], iRefStride, m_lumaClpRng.bd, COMPONENT_Y, 0, 1, m_pcEncCfg->getUseHADME() && bAllowUseOfHadamard );

  const Mv* pcMvRefine = (iFrac == 2 ? s_acMvRefineH : s_acMvRefineQ);""", """  m_pcRdCost->setDistParam( m_cDistParam, *pcPatternKey, m_filteredBlock[0][0][0], iRefStride, m_lumaClpRng.bd, COMPONENT_Y, 0, 1, m_pcEncCfg->getUseHADME() && bAllowUseOfHadamard );

  const Mv* pcMvRefine = (iFrac == 2 ? s_acMvRefineH : s_acMvRefineQ);
  const bool mmgpuFrac = mmglue::classicFracActive() && !m_skipFracME;
  if( mmgpuFrac && mmglue::enabled() )
  {
    mmglue::meSetup( m_cDistParam, m_pcEncCfg->getUseHADME() && bAllowUseOfHadamard );
    Mv ahead[9];
    for( uint32_t i = 0; i < 9; i++ )
    {
      const Mv rel = pcMvRefine[i] + baseRefMv;
      ahead[i] = mmglue::classicFracMv( Mv( rel.hor * iFrac, rel.ver * iFrac ) );
    }
    mmglue::mePrefetch( mmglue::classicFracBlock(), m_cDistParam, ahead, 9, MV_PRECISION_QUARTER );
  }
  else if( mmgpuFrac && mmglue::dumping() )
  {
    mmglue::meDumpSetup( m_cDistParam, m_pcEncCfg->getUseHADME() && bAllowUseOfHadamard );
  }""")
    t = sub(t, """    m_cDistParam.cur.buf   = piRefPos;
    dist                   = m_cDistParam.distFunc(m_cDistParam);
    dist += m_pcRdCost->getCostOfVectorWithPredictor(cMvTest.getHor(), cMvTest.getVer(), 0);
""", """    m_cDistParam.cur.buf   = piRefPos;
    if( mmgpuFrac )
    {
      const Mv absQ = mmglue::classicFracMv( Mv( horVal, verVal ) );
      if( mmglue::enabled() )
      {
        mmglue::meEval( mmglue::classicFracBlock(), m_cDistParam, absQ, MV_PRECISION_QUARTER );
        if( mmglue::verifying() ) mmglue::meVerify( m_cDistParam );   // cur.buf = piRefPos: the reference's up-sampled view
      }
      else mmglue::meDumpCandidate( mmglue::classicFracBlock(), absQ, MV_PRECISION_QUARTER );
    }
    dist                   = m_cDistParam.distFunc(m_cDistParam);
    dist += m_pcRdCost->getCostOfVectorWithPredictor(cMvTest.getHor(), cMvTest.getVer(), 0);
""")
    t = sub(t, """  Mv mv(_mv);
  // Luma interpolation filter with quarter pixel precision and MVReprojection requires MV_PRECISION_INTERNAL
  // for COMPONENT_Y.
""", """  Mv mv(_mv);
  if( mmglue::dumping() )
  {
    mmglue::meDumpCandidate( mmglue::MeBlock{ cuPosition, cuSize, refBuf.buf, motionModel, curPOC, refPOC, clpRng }, _mv, mvPrec );
  }
  // Luma interpolation filter with quarter pixel precision and MVReprojection requires MV_PRECISION_INTERNAL
  // for COMPONENT_Y.
""")
    return t


EDITS = {"Lib/CommonLib/InterPrediction.cpp": edit_interprediction, "Lib/CommonLib/MVReprojection.cpp": edit_mvreprojection,
         "Lib/CommonLib/Picture.cpp": edit_picture, "Lib/EncoderLib/InterSearch.cpp": edit_intersearch}


def build_patch():
    out = []
    for f in FILES:
        a = open(os.path.join(SRC, f), encoding="latin-1").read()
        b = EDITS[f](a)
        out += difflib.unified_diff(a.splitlines(True), b.splitlines(True), "a/" + f, "b/" + f, n=3)
    return out


def main():
    out = build_patch()
    with open(os.path.join(HERE, "mmgpu.patch"), "w", encoding="latin-1") as fh:
        fh.writelines(out)
    print("wrote mmgpu.patch: %d lines" % len(out))


if __name__ == "__main__":
    main()

// TEST INFRASTRUCTURE: golden-vector generator for the prediction parameters of a luma CU.  Calls the UNMODIFIED
// reference's IntraPrediction::initPredIntraParams (CL/IntraPrediction.cpp:487-618; private, reached by the access-specifier
// define below -- the compiled function out of oracle/_ref/libvtmref.a is what runs) for the 17 luma CU shapes x 67 modes x
// reference lines 0 / 1 / 3 and prints m_ipaParam.  Built and run by `make -f oracle/Makefile.ref intra_params`; the output is
// committed as tests/golden/intra_params.txt.gz.  Never linked into the product library.
#include <cstdio>
#include <sstream>
#include <vector>
#include <map>
#include <list>
#include <algorithm>
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/Slice.h"
#include "CommonLib/Picture.h"
#include "CommonLib/MatrixIntraPrediction.h"
#define private public
#define protected public
#include "CommonLib/IntraPrediction.h"
#undef private
#undef protected

int main()
{
  static SPS sps;
  static IntraPrediction ip;
  const int sizes[4] = { 4, 8, 16, 32 };
  const int mrls[3] = { 0, 1, 3 };
  printf( "# w h mode mrl | isModeVer refFilterFlag interpolationFlag applyPDPC intraPredAngle invAngle angularScale (angle fields as left by the call: stale for planar / DC, scale stale unless the angle is positive)\n" );
  for( int s = 0; s < 17; s++ )
  {
    const int w = s == 16 ? 64 : sizes[s >> 2], h = s == 16 ? 64 : sizes[s & 3];
    for( int m = 0; m < 3; m++ )
    for( int mode = 0; mode < 67; mode++ )
    {
      if( mrls[m] && mode == 0 ) continue;   // planar is never tried with a further reference line
      CodingUnit cu;
      cu.UnitArea::operator=( UnitArea( CHROMA_400, Area( 128, 128, w, h ) ) );
      cu.chromaFormat = CHROMA_400; cu.predMode = MODE_INTRA; cu.ispMode = 0; cu.bdpcmMode = 0; cu.mipFlag = false;
      PredictionUnit pu;
      pu.UnitArea::operator=( cu );
      pu.cu = &cu; pu.chromaFormat = CHROMA_400; pu.intraDir[0] = mode; pu.intraDir[1] = 0; pu.multiRefIdx = mrls[m];
      ip.m_ipaParam = IntraPrediction::IntraPredParam();
      ip.initPredIntraParams( pu, cu.Y(), sps );
      const IntraPrediction::IntraPredParam& p = ip.m_ipaParam;
      printf( "%d %d %d %d | %d %d %d %d %d %d %d\n", w, h, mode, mrls[m], p.isModeVer, p.refFilterFlag, p.interpolationFlag, p.applyPDPC, p.intraPredAngle, p.invAngle, p.angularScale );
    }
  }
  // intra sub-partition CUs: the prediction regions vvcb_isp_plan names (block sizes of getISPSplitDim, 4 wide at least for a vertical split)
  printf( "# ISP: cu_w cu_h isp_mode pred_w pred_h mode | the same fields\n" );
  const int all[5] = { 4, 8, 16, 32, 64 };
  for( int wi = 0; wi < 5; wi++ )
  for( int hi = 0; hi < 5; hi++ )
  for( int split = 1; split <= 2; split++ )
  {
    const int w = all[wi], h = all[hi];
    if( !CU::canUseISP( w, h, 64 ) ) continue;
    const int dim = (int) CU::getISPSplitDim( w, h, split == 1 ? TU_1D_HORZ_SPLIT : TU_1D_VERT_SPLIT );
    CompArea reg( COMPONENT_Y, CHROMA_400, 128, 128, split == 2 ? dim : w, split == 1 ? dim : h );
    if( split == 2 && CU::isMinWidthPredEnabledForBlkSize( w, h ) ) CU::adjustPredArea( reg );
    for( int mode = 0; mode < 67; mode++ )
    {
      CodingUnit cu;
      cu.UnitArea::operator=( UnitArea( CHROMA_400, Area( 128, 128, w, h ) ) );
      cu.chromaFormat = CHROMA_400; cu.predMode = MODE_INTRA; cu.ispMode = split; cu.bdpcmMode = 0; cu.mipFlag = false;
      PredictionUnit pu;
      pu.UnitArea::operator=( cu );
      pu.cu = &cu; pu.chromaFormat = CHROMA_400; pu.intraDir[0] = mode; pu.intraDir[1] = 0; pu.multiRefIdx = 0;
      ip.m_ipaParam = IntraPrediction::IntraPredParam();
      ip.initPredIntraParams( pu, reg, sps );
      const IntraPrediction::IntraPredParam& p = ip.m_ipaParam;
      printf( "ISP %d %d %d %d %d %d | %d %d %d %d %d %d %d\n", w, h, split, reg.width, reg.height, mode, p.isModeVer, p.refFilterFlag, p.interpolationFlag, p.applyPDPC, p.intraPredAngle, p.invAngle, p.angularScale );
    }
  }
  return 0;
}

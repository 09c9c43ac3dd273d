// TEST INFRASTRUCTURE: golden-vector generator for the intra sub-partition (ISP) planner (vvcb_isp_plan).
// Calls the UNMODIFIED reference's own functions out of oracle/_ref/libvtmref.a and prints one line per case:
//   CU::canUseISP            CL/UnitTools.cpp:426
//   CU::getISPSplitDim       CL/UnitTools.cpp:437
//   CU::isMinWidthPredEnabledForBlkSize / adjustPredArea   CL/UnitTools.cpp:4342-4355
//   TrQuant::getTrTypes      CL/TrQuant.cpp:752 (the ISP / implicit branch, on a hand-made luma TU of an intra CU)
// Built and run by `make -f oracle/Makefile.ref isp_geometry` in the container that has /root/reference; the
// output is committed as tests/golden/isp_geometry.txt.  Never linked into the product library.
#include <cstdio>
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/TrQuant.h"
#include "CommonLib/Slice.h"

int main()
{
  static SPS sps;
  CodingStructure cs( g_globalUnitCache.cuCache, g_globalUnitCache.puCache, g_globalUnitCache.tuCache );
  cs.sps = &sps;
  TrQuant trq;
  const int sizes[5] = { 4, 8, 16, 32, 64 };
  printf( "# w h max_tb split(1=hor,2=ver) allowed part_size n_parts min_width_pred | per part: x y w h tr_hor tr_ver (mts on) tr_hor tr_ver (mts off)\n" );
  for( int maxTb = 32; maxTb <= 64; maxTb <<= 1 )
  for( int wi = 0; wi < 5; wi++ )
  for( int hi = 0; hi < 5; hi++ )
  for( int split = 1; split <= 2; split++ )
  {
    const int w = sizes[wi], h = sizes[hi];
    const bool ok = CU::canUseISP( w, h, maxTb );
    printf( "%d %d %d %d %d", w, h, maxTb, split, ok ? 1 : 0 );
    if( !ok ) { printf( "\n" ); continue; }
    const PartSplit ps = split == 1 ? TU_1D_HORZ_SPLIT : TU_1D_VERT_SPLIT;
    const int dim = (int) CU::getISPSplitDim( w, h, ps );
    const int n   = ( split == 1 ? h : w ) >> floorLog2( dim );
    printf( " %d %d %d |", dim, n, ( split == 2 && CU::isMinWidthPredEnabledForBlkSize( w, h ) ) ? 1 : 0 );
    for( int i = 0; i < n; i++ )
    {
      const int x = split == 2 ? i * dim : 0, y = split == 1 ? i * dim : 0;
      const int tw = split == 2 ? dim : w, th = split == 1 ? dim : h;
      CodingUnit cu;
      cu.predMode = MODE_INTRA; cu.ispMode = split; cu.lfnstIdx = 0; cu.mipFlag = false; cu.sbtInfo = 0; cu.cs = &cs;
      TransformUnit tu;
      tu.cu = &cu; tu.cs = &cs; tu.chromaFormat = CHROMA_420;
      tu.blocks.clear();
      tu.blocks.push_back( CompArea( COMPONENT_Y, CHROMA_420, 64 + x, 64 + y, tw, th ) );
      int th1, tv1, th0, tv0;
      sps.setUseMTS( true );  sps.setUseIntraMTS( true );
      trq.getTrTypes( tu, COMPONENT_Y, th1, tv1 );
      sps.setUseMTS( false ); sps.setUseIntraMTS( false );
      trq.getTrTypes( tu, COMPONENT_Y, th0, tv0 );
      printf( " %d %d %d %d %d %d %d %d", x, y, tw, th, th1, tv1, th0, tv0 );
    }
    printf( "\n" );
  }
  return 0;
}

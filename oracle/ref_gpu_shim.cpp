// TEST INFRASTRUCTURE (integration demonstrator) -- the UNMODIFIED reference encoder with its rough-mode-decision predictions
// served by libvvc_intra_b200.so.
//
// The reference has no plugin interface on this path and its sources are not ours to edit, so the binding a
// maintainer would write inside IntraSearch::estIntraPredLumaQT is attached at link time instead: `ld --wrap`
// around the cross-object calls of the RMD block (seam S3/S4).  Per visit of estIntraPredLumaQT that runs the RMD block the shim
//   1. builds the vvcb_rmd_visit exactly as INTEGRATION.md section 2 describes (availability, MPMs, context prices, lambda),
//   2. pushes the reconstructed neighbourhood of the CU (vvcb_reco_update) and asks the engine for the candidate lists
//      (vvcb_rmd_eval) and every prediction block (vvcb_rmd_pred_all),
//   3. replaces the output of every IntraPrediction::predIntraAng / predIntraMip call of the RMD block by the engine's samples
//      (after checking them against the reference's own), so the encoder's SAD/SATD, lists and everything downstream are computed
//      from GPU predictions,
//   4. on return compares the engine's RD / HAD / full-RD candidate lists (modes and IEEE-double costs) with the ones the
//      reference saved (m_uiSavedRdModeListLFNST, m_dSavedModeCostLFNST, m_uiSavedHadModeListLFNST, m_savedRdModeList).
//   5. TU coding (seam S2 / a11-a14): for every luma TU without ISP / BDPCM, TrQuant::transformNxN(trModes) -- the MTS pre-selection -- is
//      repeated by vvcb_tu_eval (coefficients of every candidate transform, the pre-selection sums, vvcb_mts_preselect) and
//      TrQuant::transformNxN(quant) by vvcb_tu_eval with the quantiser the reference runs (dependent quantisation with the context prices
//      of the estimator snapshot, or RDOQ for transform skip; LFNST included): coefficients, levels and absSum must be identical, the
//      encoder goes on with the engine's levels, and the engine's reconstruction is compared with what invTransformNxN + reconstruct
//      produce afterwards,
//   6. CABACWriter::residual_coding on the bit estimator is repeated by vvcb_residual_bits from the estimator's context states.
// Any difference aborts the encoder.  At exit a summary goes to $VVCB_SHIM_REPORT.  tests/test_gpu_parity.py runs the result
// (oracle/_ref/EncoderAppGpu, built by oracle/Makefile.ref in the container that has /root/reference) next to the plain
// oracle/_ref/EncoderApp and requires byte-identical bitstreams.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <map>
#include <list>
#include <set>
#include <array>
#include <algorithm>
#include <functional>
#include <memory>
#include <sstream>
#include <iostream>
#include <fstream>
#include <mutex>
#include <cmath>
#include <limits>
#include <deque>
#include <bitset>
#include <unordered_map>
#include <atomic>
#include <chrono>
#include <iomanip>
#include <cassert>
#include <numeric>
#include <stack>
#include <stdexcept>
#include <utility>
#include <type_traits>
#include <exception>
#include <iterator>
#include <tuple>
#include <cstdarg>
#include <cstddef>
#include <climits>

#define private public
#define protected public
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/Picture.h"
#include "CommonLib/IntraPrediction.h"
#include "CommonLib/RdCost.h"
#include "CommonLib/TrQuant.h"
#include "CommonLib/Contexts.h"
#include "CommonLib/ContextModelling.h"
#include "CommonLib/UnitPartitioner.h"
#include "EncoderLib/CABACWriter.h"
#include "EncoderLib/IntraSearch.h"
#undef private
#undef protected

#include "../include/vvc_intra_b200.h"

namespace {

vvcb_ctx*            g_gpu      = nullptr;
int                  g_poc      = -1 << 30;
bool                 g_inRmd    = false;
vvcb_rmd_visit       g_visit;
vvcb_rmd_result      g_res;
vvcb_rmd_detail      g_det;
std::vector<int16_t> g_pred;              // [VVCB_NUM_SLOTS][h][w] of the current visit
int                  g_w = 0, g_h = 0;
long                 g_visits = 0, g_preds = 0, g_lists = 0;
long                 g_tuPre = 0, g_tuPreCand = 0, g_tuQuant = 0, g_tuQuantDq = 0, g_tuQuantTs = 0, g_tuQuantLfnst = 0, g_tuReco = 0, g_tuBits = 0;
const bool           g_serveTu = !getenv( "VVCB_SHIM_NO_TU" );

// the TU whose quantisation the engine has just repeated: its prediction and the engine's reconstruction wait for invTransformNxN
struct PendingTu { const TransformUnit* tu = nullptr; int x = 0, y = 0, w = 0, h = 0; std::vector<int16_t> pred, reco; } g_pend;

void die( const char* what, const char* detail = "" )
{
  fprintf( stderr, "vvcb shim: %s %s\n", what, detail );
  fflush( stderr );
  abort();
}

void gpuCheck( int rc, const char* what )
{
  if( rc != VVCB_OK ) die( what, vvcb_last_error( g_gpu ) );
}

void report()
{
  if( const char* p = getenv( "VVCB_SHIM_REPORT" ) )
    if( FILE* f = fopen( p, "w" ) )
    {
      fprintf( f, "{\"visits\": %ld, \"predictions_replaced\": %ld, \"lists_compared\": %ld, \"tu_preselections\": %ld, \"tu_preselection_candidates\": %ld, "
                  "\"tu_quantised\": %ld, \"tu_dep_quant\": %ld, \"tu_rdoq_ts\": %ld, \"tu_lfnst\": %ld, \"tu_reconstructions\": %ld, \"tu_residual_bits\": %ld, \"mismatches\": 0}\n",
               g_visits, g_preds, g_lists, g_tuPre, g_tuPreCand, g_tuQuant, g_tuQuantDq, g_tuQuantTs, g_tuQuantLfnst, g_tuReco, g_tuBits );
      fclose( f );
    }
  if( g_gpu ) vvcb_destroy( g_gpu );
}

void ensureFrame( const CodingStructure& cs )
{
  const SPS& sps = *cs.sps;
  if( !g_gpu )
  {
    if( vvcb_create( &g_gpu, 0, sps.getBitDepth( CHANNEL_TYPE_LUMA ), sps.getMaxCUWidth() ) != VVCB_OK ) die( "vvcb_create:", vvcb_last_error( nullptr ) );
    gpuCheck( vvcb_set_option( g_gpu, VVCB_OPT_DEP_QUANT, cs.slice->getDepQuantEnabledFlag() ? 1 : 0 ), "vvcb_set_option:" );
    atexit( report );
  }
  if( cs.slice->getPOC() != g_poc )
  {
    g_poc = cs.slice->getPOC();
    const CPelBuf org = cs.picture->getOrigBuf( COMPONENT_Y );         // constant during the CTU loop (EL/EncGOP.cpp:1692)
    gpuCheck( vvcb_frame_begin( g_gpu, org.buf, org.stride, org.width, org.height ), "vvcb_frame_begin:" );
  }
}

bool unitAvail( const CodingStructure& cs, const CodingUnit& cu, const Position& p )
{
  return cs.isDecomp( p, CH_L ) && cs.getCURestricted( p, cu, CH_L ) != nullptr;
}

int slotOf( const PredictionUnit& pu, bool mip )
{
  const int mode = pu.intraDir[0];
  if( mip ) return VVCB_SLOT_MIP + mode;
  if( pu.multiRefIdx == 0 ) return mode;
  for( int i = 1; i < 6; i++ )
    if( g_visit.mpm[i] == mode ) return ( pu.multiRefIdx == 1 ? VVCB_SLOT_MRL1 : VVCB_SLOT_MRL3 ) + i - 1;
  die( "multi-reference-line mode is not one of MPM[1..5]" );
  return -1;
}

void servePrediction( PelBuf& pred, const PredictionUnit& pu, bool mip )
{
  if( (int) pred.width != g_w || (int) pred.height != g_h ) die( "prediction block size differs from the visit" );
  const int16_t* src = g_pred.data() + (size_t) slotOf( pu, mip ) * g_w * g_h;
  for( int y = 0; y < g_h; y++ )
    for( int x = 0; x < g_w; x++ )
    {
      if( pred.at( x, y ) != src[y * g_w + x] ) die( "prediction sample differs from the reference's" );
      pred.at( x, y ) = src[y * g_w + x];          // the encoder goes on with the engine's samples
    }
  g_preds++;
}

bool sameList( int n, const vvcb_mode* m, const double* c, const static_vector<IntraSearch::ModeInfo, FAST_UDI_MAX_RDMODE_NUM>& refM, const static_vector<double, FAST_UDI_MAX_RDMODE_NUM>& refC )
{
  if( n != (int) refM.size() ) return false;
  for( int i = 0; i < n; i++ )
    if( m[i].mip != refM[i].mipFlg || m[i].mrl != refM[i].mRefId || m[i].mode != refM[i].modeId || memcmp( &c[i], &refC[i], sizeof( double ) ) ) return false;
  return true;
}

} // namespace

extern "C" {

bool __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib( IntraSearch*, CodingUnit&, Partitioner&, double, bool, int, int, bool );
void __real__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction*, ComponentID, PelBuf&, const PredictionUnit& );
void __real__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction*, ComponentID, PelBuf&, const PredictionUnit& );
void __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi( TrQuant*, TransformUnit&, const ComponentID&, const QpParam&, std::vector<TrMode>*, int );
void __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb( TrQuant*, TransformUnit&, const ComponentID&, const QpParam&, TCoeff&, const Ctx&, bool );

bool __wrap__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib( IntraSearch* is, CodingUnit& cu, Partitioner& pm, double best, bool a, int b, int c, bool d )
{
  const CodingStructure& cs  = *cu.cs;
  const SPS&             sps = *cs.sps;
  const int w = pm.currArea().lwidth(), h = pm.currArea().lheight();
  const bool loadFlag = sps.getUseLFNST() && cu.lfnstIdx != 0;
  int mtsUsage = 0;
  if( w <= MTS_INTRA_MAX_CU_SIZE && h <= MTS_INTRA_MAX_CU_SIZE && sps.getUseIntraMTS() )
    mtsUsage = ( sps.getUseLFNST() && cu.mtsFlag == 1 ) ? 2 : 1;
  const bool rmdRuns = mtsUsage != 2 && !loadFlag;           // EL/IntraSearch.cpp:430

  if( rmdRuns )
  {
    ensureFrame( cs );
    PredictionUnit& pu = *cu.firstPU;
    const Position lt = pu.Y();
    vvcb_rmd_visit& v = g_visit;
    memset( &v, 0, sizeof( v ) );
    v.x = lt.x; v.y = lt.y; v.log2w = floorLog2( w ); v.log2h = floorLog2( h );
    v.avail_al = unitAvail( cs, cu, lt.offset( -1, -1 ) );
    int n;
    for( n = 0; n < w / 4 && unitAvail( cs, cu, lt.offset( 4 * n, -1 ) ); n++ ) {}
    v.n_above = n;
    for( n = 0; n < w / 4 && unitAvail( cs, cu, lt.offset( w + 4 * n, -1 ) ); n++ ) {}
    v.n_above_right = n;
    for( n = 0; n < h / 4 && unitAvail( cs, cu, lt.offset( -1, 4 * n ) ); n++ ) {}
    v.n_left = n;
    for( n = 0; n < h / 4 && unitAvail( cs, cu, lt.offset( -1, h + 4 * n ) ); n++ ) {}
    v.n_below_left = n;
    unsigned mpm[NUM_MOST_PROBABLE_MODES];
    const int savedMrl = pu.multiRefIdx;
    pu.multiRefIdx = 0;
    v.num_mpm_cand = PU::getIntraMPMs( pu, mpm );
    pu.multiRefIdx = savedMrl;
    for( int i = 0; i < NUM_MOST_PROBABLE_MODES; i++ ) v.mpm[i] = mpm[i];
    if( !sps.getUseMIP() ) v.flags |= VVCB_VISIT_NO_MIP;
    const auto& st = is->m_CABACEstimator->getCtx().m_CtxStore_Std;
    const unsigned mipCtx = DeriveCtx::CtxMipFlag( cu );
    v.rates.mip_flag[0] = st[Ctx::MipFlag( mipCtx )].estFracBits( 0 );      v.rates.mip_flag[1] = st[Ctx::MipFlag( mipCtx )].estFracBits( 1 );
    v.rates.mrl_bin0[0] = st[Ctx::MultiRefLineIdx( 0 )].estFracBits( 0 );   v.rates.mrl_bin0[1] = st[Ctx::MultiRefLineIdx( 0 )].estFracBits( 1 );
    v.rates.mrl_bin1[0] = st[Ctx::MultiRefLineIdx( 1 )].estFracBits( 0 );   v.rates.mrl_bin1[1] = st[Ctx::MultiRefLineIdx( 1 )].estFracBits( 1 );
    v.rates.isp_bin0_0  = st[Ctx::ISPMode( 0 )].estFracBits( 0 );
    v.rates.mpm_flag[0] = st[Ctx::IntraLumaMpmFlag()].estFracBits( 0 );     v.rates.mpm_flag[1] = st[Ctx::IntraLumaMpmFlag()].estFracBits( 1 );
    v.rates.planar_flag[0] = st[Ctx::IntraLumaPlanarFlag( 1 )].estFracBits( 0 ); v.rates.planar_flag[1] = st[Ctx::IntraLumaPlanarFlag( 1 )].estFracBits( 1 );
    v.sqrt_lambda = is->m_pcRdCost->getMotionLambda( cu.transQuantBypass ) * FRAC_BITS_SCALE;

    // the reconstructed neighbourhood the reference lines come from: up to 4 rows above (lines 0, 1, 3) over 2w + 4 columns and
    // up to 4 columns to the left over 2h + 4 rows, clipped to the picture
    const CPelBuf reco = cs.picture->getRecoBuf( COMPONENT_Y );
    const int pw = cs.picture->lwidth(), ph = cs.picture->lheight();
    if( lt.y >= 4 )
    {
      const int x0 = std::max( 0, lt.x - 4 ), x1 = std::min( pw, lt.x + 2 * w + 4 );
      gpuCheck( vvcb_reco_update( g_gpu, reco.bufAt( x0, lt.y - 4 ), reco.stride, x0, lt.y - 4, x1 - x0, 4 ), "vvcb_reco_update:" );
    }
    if( lt.x >= 4 )
    {
      const int y1 = std::min( ph, lt.y + 2 * h + 4 );
      gpuCheck( vvcb_reco_update( g_gpu, reco.bufAt( lt.x - 4, lt.y ), reco.stride, lt.x - 4, lt.y, 4, y1 - lt.y ), "vvcb_reco_update:" );
    }
    gpuCheck( vvcb_rmd_eval( g_gpu, &v, 1, &g_res, &g_det ), "vvcb_rmd_eval:" );
    g_pred.resize( (size_t) VVCB_NUM_SLOTS * w * h );
    gpuCheck( vvcb_rmd_pred_all( g_gpu, &v, g_pred.data() ), "vvcb_rmd_pred_all:" );
    g_w = w; g_h = h;
    g_inRmd = true;
    g_visits++;
  }

  const bool ret = __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib( is, cu, pm, best, a, b, c, d );

  if( rmdRuns )
  {
    g_inRmd = false;
    const bool testMip = sps.getUseMIP() && mipModesAvailable( Size( w, h ) );
    const bool early   = testMip && !allowLfnstWithMip( Size( w, h ) );       // EL/IntraSearch.cpp:686 saved the regular-only list
    bool ok;
    if( !early )
      ok = sameList( g_res.n_rd, g_res.rd_mode, g_res.rd_cost, is->m_uiSavedRdModeListLFNST, is->m_dSavedModeCostLFNST ) &&
           sameList( g_res.n_had, g_res.had_mode, g_res.had_cost, is->m_uiSavedHadModeListLFNST, is->m_dSavedHadListLFNST );
    else
    {
      const int k = (int) is->m_uiSavedRdModeListLFNST.size(), kh = (int) is->m_uiSavedHadModeListLFNST.size();
      ok = k <= g_det.n_reg && kh <= g_det.n_reg_had &&
           sameList( k, g_det.reg_mode, g_det.reg_cost, is->m_uiSavedRdModeListLFNST, is->m_dSavedModeCostLFNST ) &&
           sameList( kh, g_det.reg_had_mode, g_det.reg_had_cost, is->m_uiSavedHadModeListLFNST, is->m_dSavedHadListLFNST );
    }
    if( ok && mtsUsage == 1 )                                                 // the full-RD list saved at :884-889
    {
      ok = g_res.n_final == is->m_savedNumRdModes[0];
      for( int i = 0; ok && i < g_res.n_final; i++ )
      {
        const auto& m = is->m_savedRdModeList[0][i];
        ok = g_res.final_mode[i].mip == m.mipFlg && g_res.final_mode[i].mrl == m.mRefId && g_res.final_mode[i].mode == m.modeId;
      }
    }
    if( !ok ) die( "candidate lists differ from the reference's" );
    g_lists++;
  }
  return ret;
}

void __wrap__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction* ip, ComponentID c, PelBuf& pred, const PredictionUnit& pu )
{
  __real__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit( ip, c, pred, pu );
  if( g_inRmd && c == COMPONENT_Y ) servePrediction( pred, pu, false );
}

void __wrap__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction* ip, ComponentID c, PelBuf& pred, const PredictionUnit& pu )
{
  __real__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit( ip, c, pred, pu );
  if( g_inRmd && c == COMPONENT_Y ) servePrediction( pred, pu, true );
}

} // extern "C"

// ---- TU coding (a11-a14) -------------------------------------------------------------------------------------------------------
namespace {

bool tuServed( const TransformUnit& tu, ComponentID c )
{
  return g_serveTu && c == COMPONENT_Y && !tu.noResidual && !tu.cu->ispMode && !tu.cu->bdpcmMode && CU::isIntra( *tu.cu );
}

void denseResidualAndPrediction( const TransformUnit& tu, std::vector<int16_t>& resi, std::vector<int16_t>& pred )
{
  const CompArea& rect = tu.blocks[COMPONENT_Y];
  const CPelBuf r = tu.cs->getResiBuf( rect );
  const CPelBuf o = tu.cs->picture->getOrigBuf( COMPONENT_Y );
  const int w = rect.width, h = rect.height;
  resi.resize( w * h ); pred.resize( w * h );
  for( int y = 0; y < h; y++ )
    for( int x = 0; x < w; x++ )
    {
      resi[y * w + x] = r.at( x, y );
      pred[y * w + x] = o.at( rect.x + x, rect.y + y ) - r.at( x, y );     // resi = org - pred (EL/IntraSearch.cpp:2922)
    }
}

void fillJobGeometry( vvcb_tu_job& j, const TransformUnit& tu )
{
  const CompArea& rect = tu.blocks[COMPONENT_Y];
  memset( &j, 0, sizeof( j ) );
  j.x = rect.x; j.y = rect.y; j.log2w = floorLog2( rect.width ); j.log2h = floorLog2( rect.height );
}

void fillRates( vvcb_dq_rates& out, const Ctx& ctx )
{
  const FracBitsAccess& fb = ctx.getFracBitsAcess();
  uint32_t* p = reinterpret_cast<uint32_t*>( &out );
  auto put = [&]( const CtxSet& set, int num ) { for( int i = 0; i < num; i++ ) { const BinFracBits b = fb.getFracBitsArray( set( i ) ); *p++ = b.intBits[0]; *p++ = b.intBits[1]; } };
  put( Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2 );
  for( int st = 0; st < 3; st++ ) put( Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * st], 12 );
  put( Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21 );
  put( Ctx::LastX[CHANNEL_TYPE_LUMA], 20 ); put( Ctx::LastY[CHANNEL_TYPE_LUMA], 20 );
  put( Ctx::TsSigCoeffGroup, 3 ); put( Ctx::TsSigFlag, 3 ); put( Ctx::TsParFlag, 1 ); put( Ctx::TsGtxFlag, 5 ); put( Ctx::TsLrg1Flag, 4 ); put( Ctx::TsResidualSign, 6 );
  if( (char*) p != (char*) &out + sizeof( out ) ) die( "vvcb_dq_rates layout" );
}

void fillStates( vvcb_ctx_states& out, const Ctx& ctx )
{
  vvcb_bin_model* p = reinterpret_cast<vvcb_bin_model*>( &out );
  auto put = [&]( const CtxSet& set, int num ) { for( int i = 0; i < num; i++ ) { const BinProbModel_Std& m = ctx.m_CtxStore_Std[set( i )]; p->state[0] = m.m_state[0]; p->state[1] = m.m_state[1]; p->rate = m.m_rate; p->pad = 0; p++; } };
  put( Ctx::MTSIndex, 11 );
  put( Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2 );
  for( int k = 0; k < 3; k++ ) put( Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * k], 12 );
  put( Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21 );
  put( Ctx::LastX[CHANNEL_TYPE_LUMA], 20 ); put( Ctx::LastY[CHANNEL_TYPE_LUMA], 20 );
  put( Ctx::TsSigCoeffGroup, 3 ); put( Ctx::TsSigFlag, 3 ); put( Ctx::TsParFlag, 1 ); put( Ctx::TsGtxFlag, 5 ); put( Ctx::TsLrg1Flag, 4 ); put( Ctx::TsResidualSign, 6 );
  if( (char*) p != (char*) &out + sizeof( out ) ) die( "vvcb_ctx_states layout" );
}

} // namespace

extern "C" {

void __real__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam( TrQuant*, TransformUnit&, const ComponentID&, PelBuf&, const QpParam& );
void __real__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx( CABACWriter*, const TransformUnit&, ComponentID, CUCtx* );

// TrQuant::transformNxN(trModes), CL/TrQuant.cpp:1049-1124: every candidate transform of the TU + the pre-selection.  (The RMD block of a
// visit ends where its first TU is coded.)
void __wrap__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi( TrQuant* tq, TransformUnit& tu, const ComponentID& c, const QpParam& qp, std::vector<TrMode>* modes, int maxCand )
{
  g_inRmd = false;
  const bool serve = tuServed( tu, c ) && tu.cu->lfnstIdx == 0 && !modes->empty();
  std::vector<int16_t> resi, pred;
  if( serve ) { ensureFrame( *tu.cs ); denseResidualAndPrediction( tu, resi, pred ); }
  __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi( tq, tu, c, qp, modes, maxCand );
  if( !serve ) return;
  const CompArea& rect = tu.blocks[c];
  const int n = rect.width * rect.height, k = (int) modes->size();
  std::vector<vvcb_tu_job> jobs( k );
  std::vector<int16_t> resiAll( (size_t) n * k );
  for( int i = 0; i < k; i++ )
  {
    fillJobGeometry( jobs[i], tu );
    jobs[i].mts_idx = ( *modes )[i].first;
    jobs[i].offset  = i * n;
    memcpy( &resiAll[(size_t) i * n], resi.data(), n * sizeof( int16_t ) );
  }
  std::vector<int32_t> coeff( (size_t) n * k );
  std::vector<vvcb_tu_result> res( k );
  gpuCheck( vvcb_tu_eval( g_gpu, jobs.data(), k, resiAll.data(), nullptr, (size_t) n * k, nullptr, nullptr, 0, coeff.data(), nullptr, nullptr, res.data() ), "vvcb_tu_eval (pre-selection):" );
  std::vector<int32_t> sums( k );
  std::vector<uint8_t> sel( k );
  for( int i = 0; i < k; i++ )
  {
    sums[i] = res[i].abs_sum_coeff;
    if( memcmp( &coeff[(size_t) i * n], tq->m_mtsCoeffs[( *modes )[i].first], n * sizeof( int32_t ) ) ) die( "coefficients of a candidate transform differ from the reference's" );
  }
  vvcb_mts_preselect( sums.data(), k, rect.width, rect.height, maxCand, sel.data() );
  for( int i = 0; i < k; i++ )
    if( ( sel[i] != 0 ) != ( *modes )[i].second ) die( "MTS pre-selection differs from the reference's" );
  g_tuPre++; g_tuPreCand += k;
}

// TrQuant::transformNxN(quant), CL/TrQuant.cpp:1127-1260: (forward transform, LFNST,) quantisation with the estimator's context prices
void __wrap__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb( TrQuant* tq, TransformUnit& tu, const ComponentID& c, const QpParam& qp, TCoeff& absSum, const Ctx& ctx, bool loadTr )
{
  g_inRmd = false;
  g_pend.tu = nullptr;
  const bool ts = tu.mtsIdx == MTS_SKIP;
  const bool dq = tu.cs->slice->getDepQuantEnabledFlag() && !ts;
  const bool serve = tuServed( tu, c ) && ( dq || ( ts && tq->m_quant->m_useRDOQ && tq->m_quant->m_useRDOQTS ) );
  std::vector<int16_t> resi, pred;
  if( serve ) { ensureFrame( *tu.cs ); denseResidualAndPrediction( tu, resi, pred ); }
  __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb( tq, tu, c, qp, absSum, ctx, loadTr );
  if( !serve ) return;
  const CompArea& rect = tu.blocks[c];
  const int w = rect.width, h = rect.height, n = w * h;
  vvcb_tu_job j;
  fillJobGeometry( j, tu );
  j.mts_idx = tu.mtsIdx;
  j.flags   = VVCB_TU_QUANT | ( dq ? VVCB_TU_DEPQUANT : VVCB_TU_RDOQ_TS );
  j.qp_per  = qp.per( ts ); j.qp_rem = qp.rem( ts );
  j.lfnst_idx = ts ? 0 : tu.cu->lfnstIdx;
  if( j.lfnst_idx )
  {
    const PredictionUnit& pu = *tu.cs->getPU( rect.pos(), CHANNEL_TYPE_LUMA );
    j.intra_mode = PU::isMIP( pu, CHANNEL_TYPE_LUMA ) ? PLANAR_IDX : (int) PU::getFinalIntraMode( pu, CHANNEL_TYPE_LUMA );
  }
  const BinFracBits cbf = ctx.getFracBitsAcess().getFracBitsArray( Ctx::QtCbf[COMPONENT_Y]( DeriveCtx::CtxQtCbf( COMPONENT_Y, tu.cbf[COMPONENT_Cb] ) ) );
  j.cbf_delta_bits = int32_t( cbf.intBits[1] ) - int32_t( cbf.intBits[0] );      // RateEstimator::xSetLastCoeffOffset, CL/DepQuant.cpp:531-540
  j.lambda = tq->m_quant->getLambda();
  static vvcb_dq_rates rates;
  fillRates( rates, ctx );
  std::vector<int32_t> coeff( n ), level( n );
  g_pend.reco.resize( n );
  vvcb_tu_result res;
  gpuCheck( vvcb_tu_eval( g_gpu, &j, 1, resi.data(), pred.data(), n, &rates, nullptr, 1, coeff.data(), level.data(), g_pend.reco.data(), &res ), "vvcb_tu_eval (quantisation):" );
  // coefficients: with LFNST the reference's buffer keeps stale values outside the top-left 8x8 / 4x4 (tests/test_oracle_lfnst.py)
  const TCoeff* co = loadTr ? tq->m_mtsCoeffs[tu.mtsIdx] : tq->m_tempCoeff;
  const int sb = j.lfnst_idx ? ( std::min( w, h ) >= 8 ? 8 : 4 ) : 1 << 30;
  for( int y = 0; y < h; y++ )
    for( int x = 0; x < w; x++ )
      if( x < sb && y < sb ? coeff[y * w + x] != co[y * w + x] : ( j.lfnst_idx && coeff[y * w + x] != 0 ) ) die( "transform coefficients differ from the reference's" );
  CoeffBuf lv = tu.getCoeffs( c );
  for( int y = 0; y < h; y++ )
    for( int x = 0; x < w; x++ )
    {
      if( lv.at( x, y ) != level[y * w + x] ) die( dq ? "dependent-quantisation levels differ from the reference's" : "transform-skip RDOQ levels differ from the reference's" );
      lv.at( x, y ) = level[y * w + x];            // the encoder goes on with the engine's levels
    }
  if( res.abs_sum_level != absSum ) die( "absSum differs from the reference's" );
  absSum = res.abs_sum_level;
  g_tuQuant++; g_tuQuantDq += dq; g_tuQuantTs += !dq; g_tuQuantLfnst += j.lfnst_idx != 0;
  if( absSum > 0 ) { g_pend.tu = &tu; g_pend.x = rect.x; g_pend.y = rect.y; g_pend.w = w; g_pend.h = h; g_pend.pred.swap( pred ); }
}

// TrQuant::invTransformNxN, CL/TrQuant.cpp:561 (+ PelBuf::reconstruct, EL/IntraSearch.cpp:3050): the engine's reconstruction of the TU it
// has just quantised must equal clip( pred + inverse residual )
void __wrap__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam( TrQuant* tq, TransformUnit& tu, const ComponentID& c, PelBuf& resi, const QpParam& qp )
{
  __real__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam( tq, tu, c, resi, qp );
  if( c != COMPONENT_Y || g_pend.tu != &tu ) return;
  const CompArea& rect = tu.blocks[c];
  if( rect.x != g_pend.x || rect.y != g_pend.y || (int) rect.width != g_pend.w || (int) rect.height != g_pend.h || (int) resi.width != g_pend.w || (int) resi.height != g_pend.h ) { g_pend.tu = nullptr; return; }
  const int maxv = ( 1 << tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) ) - 1;
  for( int y = 0; y < g_pend.h; y++ )
    for( int x = 0; x < g_pend.w; x++ )
    {
      const int v = std::min( maxv, std::max( 0, g_pend.pred[y * g_pend.w + x] + resi.at( x, y ) ) );
      if( v != g_pend.reco[y * g_pend.w + x] ) die( "reconstruction differs from the reference's" );
    }
  g_pend.tu = nullptr;
  g_tuReco++;
}

// CABACWriter::residual_coding on the bit estimator (EL/CABACWriter.cpp:3773, from IntraSearch::xEncCoeffQT)
void __wrap__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx( CABACWriter* cw, const TransformUnit& tu, ComponentID c, CUCtx* cuCtx )
{
  const bool serve = g_serveTu && g_gpu && c == COMPONENT_Y && !cw->m_BinEncoder.isEncoding() && !tu.cu->ispMode && !tu.cu->bdpcmMode && CU::isIntra( *tu.cu );
  static vvcb_ctx_states states;
  uint64_t before = 0;
  if( serve ) { fillStates( states, cw->getCtx() ); before = cw->m_BinEncoder.getEstFracBits(); }
  __real__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx( cw, tu, c, cuCtx );
  if( !serve ) return;
  const CompArea& rect = tu.blocks[c];
  const int w = rect.width, h = rect.height;
  vvcb_tu_job j;
  fillJobGeometry( j, tu );
  j.mts_idx = tu.mtsIdx;
  j.flags   = ( TU::isTSAllowed( tu, c ) ? VVCB_TU_TS_ALLOWED : 0 ) | ( TU::isMTSAllowed( tu, c ) ? VVCB_TU_MTS_ALLOWED : 0 );
  std::vector<int32_t> level( w * h );
  const CCoeffBuf lv = tu.getCoeffs( c );
  for( int y = 0; y < h; y++ ) for( int x = 0; x < w; x++ ) level[y * w + x] = lv.at( x, y );
  uint64_t bits = 0;
  gpuCheck( vvcb_residual_bits( g_gpu, &j, 1, level.data(), level.size(), &states, 1, &bits ), "vvcb_residual_bits:" );
  if( bits != cw->m_BinEncoder.getEstFracBits() - before ) die( "residual bits differ from the reference's estimator" );
  g_tuBits++;
}

} // extern "C"

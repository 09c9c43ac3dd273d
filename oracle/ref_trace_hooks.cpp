// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Link-time (--wrap) observers for the UNMODIFIED reference encoder built by
// oracle/Makefile.ref.  The reference has no golden vectors of its own
// (SURVEY.md §4), so this file turns the real encoder into a golden-vector
// generator: every call that crosses an object-file boundary on the hot path is
// intercepted with `ld --wrap=<mangled name>`, the real function is run
// untouched, and its inputs/outputs are appended to a binary trace.
//
// Wrapped call sites (all in EL/IntraSearch.cpp unless noted):
//   IntraSearch::estIntraPredLumaQT            EL/EncCu.cpp:2525          visit begin/end, final lists
//   IntraPrediction::initIntraPatternChType    :483,:645,:712             reference arrays (a1-a3)
//   IntraPrediction::predIntraAng              :511,:598,:660             prediction samples (a4,a5)
//   IntraPrediction::predIntraMip              :719                       MIP prediction (a6)
//   CABACWriter::intra_luma_pred_mode          :4263 xFracModeBitsIntra   mode bits (a10)
//   TrQuant::transformNxN (both overloads)     :2965,:2968                fwd transform / quant (a12,a13)
//   TrQuant::invTransformNxN                   :3002                      dequant + inverse (a14)
//   EncCu::updateCtuDataISlice                 EL/EncSlice.cpp:1291       per-CTU Hadamard texture sum (a16; needs --RateControl=1)
//
// Record stream: { u8 tag; u32 payload_bytes; payload }.  Layouts are parsed by
// tools/make_golden.py (one struct format per tag, kept next to each emit()).
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
#include <cmath>

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
#include "EncoderLib/EncCu.h"
#undef private
#undef protected

namespace {

FILE*  g_out         = nullptr;
bool   g_init        = false;
int    g_visitStride = 1;     // keep every Nth RMD visit per shape ...
int    g_visitFirst  = 4;     // ... after always keeping the first K of each shape
int    g_tuStride    = 64;    // keep every Nth transform call per shape
int    g_tuFirst     = 2;
int    g_fullPred    = 0;     // store prediction samples (not only their hash) for the first N visits of each shape
bool   g_curFull     = false;
int    g_maxVisits   = 1 << 30;
const char* g_only   = nullptr;  // VVC_TRACE_ONLY: keep only the records with these tags

IntraSearch* g_is      = nullptr;
bool         g_inRmd   = false;   // inside the SATD rough-mode-decision part of a recorded visit
bool         g_inVisit = false;
uint32_t     g_visitId = 0;
uint32_t     g_kept    = 0;
std::map<int, int> g_shapeCount, g_tuShapeCount;

void init()
{
  if( g_init ) return;
  g_init = true;
  const char* p = getenv( "VVC_TRACE_OUT" );
  if( !p ) return;
  g_out = fopen( p, "wb" );
  if( const char* s = getenv( "VVC_TRACE_VISIT_STRIDE" ) ) g_visitStride = std::max( 1, atoi( s ) );
  if( const char* s = getenv( "VVC_TRACE_VISIT_FIRST"  ) ) g_visitFirst  = atoi( s );
  if( const char* s = getenv( "VVC_TRACE_TU_STRIDE"    ) ) g_tuStride    = std::max( 1, atoi( s ) );
  if( const char* s = getenv( "VVC_TRACE_TU_FIRST"     ) ) g_tuFirst     = atoi( s );
  if( const char* s = getenv( "VVC_TRACE_FULL_PRED"    ) ) g_fullPred    = atoi( s );
  if( const char* s = getenv( "VVC_TRACE_MAX_VISITS"   ) ) g_maxVisits   = atoi( s );
  g_only = getenv( "VVC_TRACE_ONLY" );
}

struct Rec
{
  std::vector<uint8_t> b;
  template<class T> void put( T v ) { const uint8_t* p = (const uint8_t*) &v; b.insert( b.end(), p, p + sizeof( T ) ); }
  void i32( int v )       { put<int32_t>( v ); }
  void u32( uint32_t v )  { put<uint32_t>( v ); }
  void u64( uint64_t v )  { put<uint64_t>( v ); }
  void f64( double v )    { put<double>( v ); }
  void i16( int v )       { put<int16_t>( (int16_t) v ); }
  void emit( char tag )
  {
    if( !g_out || ( g_only && !strchr( g_only, tag ) ) ) return;
    uint8_t  t = (uint8_t) tag;
    uint32_t n = (uint32_t) b.size();
    fwrite( &t, 1, 1, g_out );
    fwrite( &n, 4, 1, g_out );
    if( n ) fwrite( b.data(), 1, n, g_out );
  }
};

uint64_t fnv1a( const int16_t* p, size_t n )
{
  uint64_t h = 1469598103934665603ull;
  const uint8_t* q = (const uint8_t*) p;
  for( size_t i = 0; i < 2 * n; i++ ) { h ^= q[i]; h *= 1099511628211ull; }
  return h;
}

void putBlock( Rec& r, const CPelBuf& b )
{
  for( int y = 0; y < (int) b.height; y++ )
    for( int x = 0; x < (int) b.width; x++ )
      r.i16( b.at( x, y ) );
}

std::vector<int16_t> flat( const CPelBuf& b )
{
  std::vector<int16_t> v;
  v.reserve( b.width * b.height );
  for( int y = 0; y < (int) b.height; y++ )
    for( int x = 0; x < (int) b.width; x++ )
      v.push_back( b.at( x, y ) );
  return v;
}

bool unitAvail( const CodingStructure& cs, const CodingUnit& cu, const Position& p )
{
  return cs.isDecomp( p, CH_L ) && cs.getCURestricted( p, cu, CH_L ) != nullptr;
}

uint64_t refDist( const CPelBuf& org, const CPelBuf& cur, int bitDepth, bool had )
{
  DistParam dp;
  g_is->m_pcRdCost->setDistParam( dp, org, cur, bitDepth, COMPONENT_Y, had );
  dp.applyWeight = false;
  return dp.distFunc( dp );
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
extern "C" {

bool __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib( IntraSearch*, CodingUnit&, Partitioner&, double, bool, int, int, bool );
void __real__ZN15IntraPrediction22initIntraPatternChTypeERK10CodingUnitRK8CompAreab( IntraPrediction*, const CodingUnit&, const CompArea&, bool );
void __real__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction*, ComponentID, PelBuf&, const PredictionUnit& );
void __real__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction*, ComponentID, PelBuf&, const PredictionUnit& );
void __real__ZN11CABACWriter20intra_luma_pred_modeERK14PredictionUnit( CABACWriter*, const PredictionUnit& );
void __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi( TrQuant*, TransformUnit&, const ComponentID&, const QpParam&, std::vector<TrMode>*, int );
void __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb( TrQuant*, TransformUnit&, const ComponentID&, const QpParam&, TCoeff&, const Ctx&, bool );
void __real__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam( TrQuant*, TransformUnit&, const ComponentID&, PelBuf&, const QpParam& );
void __real__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx( CABACWriter*, const TransformUnit&, ComponentID, CUCtx* );

// ---- visit ---------------------------------------------------------------------------------------------
// 'V': u32 visitId, i32 poc,x,y,w,h,lfnstIdx,mtsFlag,bitDepth,qp, f64 sqrtLambda, i32 mpm[6], i32 numCandMpm,
//      i32 mipCtx, u32 mipFlag[2], mrl0[2], mrl1[2], isp0, mpmFlag[2], planarFlag[2], i32 picW, picH, org[w*h] i16
// 'L': u32 visitId, i32 variant(0: list after MIP+reduce, 1: regular-only list truncated), i32 nRd, {i32 mip,mrl,mode; f64 cost}*,
//      i32 nHad, {i32 mip,mrl,mode; f64 cost}*, i32 nFinal, {i32 mip,mrl,isp,mode}*
bool __wrap__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib( IntraSearch* is, CodingUnit& cu, Partitioner& pm, double best, bool a, int b, int c, bool d )
{
  init();
  if( !g_out )
    return __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib( is, cu, pm, best, a, b, c, d );

  const CodingStructure& cs  = *cu.cs;
  const SPS&             sps = *cs.sps;
  const int w = pm.currArea().lwidth(), h = pm.currArea().lheight();
  const bool loadFlag = sps.getUseLFNST() && cu.lfnstIdx != 0;
  int mtsUsage = 0;
  if( w <= MTS_INTRA_MAX_CU_SIZE && h <= MTS_INTRA_MAX_CU_SIZE && sps.getUseIntraMTS() )
    mtsUsage = ( sps.getUseLFNST() && cu.mtsFlag == 1 ) ? 2 : 1;
  const bool rmdRuns = mtsUsage != 2 && !loadFlag;

  bool keep = false;
  if( rmdRuns && (int) g_kept < g_maxVisits )
  {
    int& n = g_shapeCount[w * 256 + h];
    keep   = n < g_visitFirst || ( n % g_visitStride ) == 0;
    g_curFull = n < g_fullPred;
    n++;
  }
  g_is = is;
  g_visitId++;
  g_inVisit = keep;
  g_inRmd   = keep;
  if( keep )
  {
    g_kept++;
    PredictionUnit& pu = *cu.firstPU;
    Rec r;
    r.u32( g_visitId );
    r.i32( cs.slice->getPOC() ); r.i32( pu.Y().x ); r.i32( pu.Y().y ); r.i32( w ); r.i32( h );
    r.i32( cu.lfnstIdx ); r.i32( cu.mtsFlag ); r.i32( sps.getBitDepth( CHANNEL_TYPE_LUMA ) ); r.i32( cu.qp );
    r.f64( is->m_pcRdCost->getMotionLambda( cu.transQuantBypass ) * FRAC_BITS_SCALE );
    unsigned mpm[NUM_MOST_PROBABLE_MODES];
    const int savedMrl = pu.multiRefIdx;
    pu.multiRefIdx = 0;
    const int numCand = PU::getIntraMPMs( pu, mpm );
    pu.multiRefIdx = savedMrl;
    for( int i = 0; i < NUM_MOST_PROBABLE_MODES; i++ ) r.i32( mpm[i] );
    r.i32( numCand );
    const Ctx& ctx = is->m_CABACEstimator->getCtx();
    const auto& st = ctx.m_CtxStore_Std;
    const unsigned mipCtx = DeriveCtx::CtxMipFlag( cu );
    r.i32( mipCtx );
    r.u32( st[Ctx::MipFlag( mipCtx )].estFracBits( 0 ) ); r.u32( st[Ctx::MipFlag( mipCtx )].estFracBits( 1 ) );
    r.u32( st[Ctx::MultiRefLineIdx( 0 )].estFracBits( 0 ) ); r.u32( st[Ctx::MultiRefLineIdx( 0 )].estFracBits( 1 ) );
    r.u32( st[Ctx::MultiRefLineIdx( 1 )].estFracBits( 0 ) ); r.u32( st[Ctx::MultiRefLineIdx( 1 )].estFracBits( 1 ) );
    r.u32( st[Ctx::ISPMode( 0 )].estFracBits( 0 ) );
    r.u32( st[Ctx::IntraLumaMpmFlag()].estFracBits( 0 ) ); r.u32( st[Ctx::IntraLumaMpmFlag()].estFracBits( 1 ) );
    r.u32( st[Ctx::IntraLumaPlanarFlag( 1 )].estFracBits( 0 ) ); r.u32( st[Ctx::IntraLumaPlanarFlag( 1 )].estFracBits( 1 ) );
    r.i32( cs.picture->lwidth() ); r.i32( cs.picture->lheight() );
    putBlock( r, cs.getOrgBuf( pu.Y() ) );
    r.emit( 'V' );
  }

  const bool ret = __real__ZN11IntraSearch18estIntraPredLumaQTER10CodingUnitR11Partitionerdbiib( is, cu, pm, best, a, b, c, d );

  if( keep )
  {
    Rec r;
    r.u32( g_visitId );
    const bool testMip = sps.getUseMIP() && mipModesAvailable( Size( w, h ) );
    const bool early   = testMip && !allowLfnstWithMip( Size( w, h ) );   // IntraSearch.cpp:686 saved the regular-only list
    r.i32( early ? 1 : 0 );
    r.i32( (int) is->m_uiSavedRdModeListLFNST.size() );
    for( size_t i = 0; i < is->m_uiSavedRdModeListLFNST.size(); i++ )
    {
      const auto& m = is->m_uiSavedRdModeListLFNST[i];
      r.i32( m.mipFlg ); r.i32( m.mRefId ); r.i32( m.modeId ); r.f64( is->m_dSavedModeCostLFNST[i] );
    }
    r.i32( (int) is->m_uiSavedHadModeListLFNST.size() );
    for( size_t i = 0; i < is->m_uiSavedHadModeListLFNST.size(); i++ )
    {
      const auto& m = is->m_uiSavedHadModeListLFNST[i];
      r.i32( m.mipFlg ); r.i32( m.mRefId ); r.i32( m.modeId ); r.f64( is->m_dSavedHadListLFNST[i] );
    }
    const int nFinal = mtsUsage == 1 ? is->m_savedNumRdModes[0] : 0;   // :884-889 (only when w,h <= 32)
    r.i32( nFinal );
    for( int i = 0; i < nFinal; i++ )
    {
      const auto& m = is->m_savedRdModeList[0][i];
      r.i32( m.mipFlg ); r.i32( m.mRefId ); r.i32( m.ispMod ); r.i32( m.modeId );
    }
    r.emit( 'L' );
  }
  g_inVisit = false;
  g_inRmd   = false;
  return ret;
}

// ---- reference samples ---------------------------------------------------------------------------------
// 'R': u32 visitId, i32 mrl, force, w, h, availAL, nAbove, nAboveRight, nLeft, nBelowLeft (units of 4),
//      i32 hasFiltered, unfTop[2w+1+mrl], unfLeft[2h+1+mrl], (filtTop, filtLeft), recoTop[4][2w+8], recoLeft[2h+4][4]
//      recoTop row r (0..3) = picture row y-4+r, columns x-4 .. x+2w+3 ; recoLeft row j = picture row y+j, columns x-4..x-1
void __wrap__ZN15IntraPrediction22initIntraPatternChTypeERK10CodingUnitRK8CompAreab( IntraPrediction* ip, const CodingUnit& cu, const CompArea& area, bool force )
{
  __real__ZN15IntraPrediction22initIntraPatternChTypeERK10CodingUnitRK8CompAreab( ip, cu, area, force );
  if( !g_out || !g_inRmd || area.compID != COMPONENT_Y ) return;

  const CodingStructure& cs = *cu.cs;
  const int mrl = cu.firstPU->multiRefIdx;
  const int w = area.width, h = area.height;
  const int stride = ip->m_topRefLength + 1 + mrl;
  Rec r;
  r.u32( g_visitId ); r.i32( mrl ); r.i32( force ); r.i32( w ); r.i32( h );
  const Position lt = area;
  r.i32( unitAvail( cs, cu, lt.offset( -1, -1 ) ) );
  int n = 0;
  for( n = 0; n < w / 4 && unitAvail( cs, cu, lt.offset( 4 * n, -1 ) ); n++ ) {}
  r.i32( n );
  for( n = 0; n < w / 4 && unitAvail( cs, cu, lt.offset( w + 4 * n, -1 ) ); n++ ) {}
  r.i32( n );
  for( n = 0; n < h / 4 && unitAvail( cs, cu, lt.offset( -1, 4 * n ) ); n++ ) {}
  r.i32( n );
  for( n = 0; n < h / 4 && unitAvail( cs, cu, lt.offset( -1, h + 4 * n ) ); n++ ) {}
  r.i32( n );
  const bool hasFilt = force || ip->m_ipaParam.refFilterFlag;
  r.i32( hasFilt );
  for( int pass = 0; pass < ( hasFilt ? 2 : 1 ); pass++ )
  {
    const Pel* p = ip->m_piYuvExt[COMPONENT_Y][pass];
    for( int i = 0; i <= 2 * w + mrl; i++ ) r.i16( p[i] );
    for( int i = 0; i <= 2 * h + mrl; i++ ) r.i16( p[i * stride] );
  }
  const CPelBuf reco = cs.picture->getRecoBuf( COMPONENT_Y );
  const int pw = reco.width, ph = reco.height;
  auto at = [&]( int x, int y ) { return reco.at( std::min( std::max( x, 0 ), pw - 1 ), std::min( std::max( y, 0 ), ph - 1 ) ); };
  for( int rr = 0; rr < 4; rr++ )
    for( int xx = 0; xx < 2 * w + 8; xx++ ) r.i16( at( area.x - 4 + xx, area.y - 4 + rr ) );
  for( int j = 0; j < 2 * h + 4; j++ )
    for( int xx = 0; xx < 4; xx++ ) r.i16( at( area.x - 4 + xx, area.y + j ) );
  r.emit( 'R' );
}

// ---- prediction ----------------------------------------------------------------------------------------
// 'P': u32 visitId, i32 mip, mode, mrl, w, h, isModeVer, refFilterFlag, interpolationFlag, applyPDPC, intraPredAngle,
//      invAngle, angularScale, u64 sad, u64 satd, u64 predHash, i32 hasSamples, (pred[w*h] i16)
static void emitPred( IntraPrediction* ip, const PelBuf& pred, const PredictionUnit& pu, bool mip )
{
  const CodingStructure& cs = *pu.cs;
  const int bd = cs.sps->getBitDepth( CHANNEL_TYPE_LUMA );
  Rec r;
  r.u32( g_visitId ); r.i32( mip ); r.i32( pu.intraDir[0] ); r.i32( mip ? 0 : pu.multiRefIdx ); r.i32( pred.width ); r.i32( pred.height );
  const auto& q = ip->m_ipaParam;
  r.i32( q.isModeVer ); r.i32( q.refFilterFlag ); r.i32( q.interpolationFlag ); r.i32( q.applyPDPC );
  r.i32( q.intraPredAngle ); r.i32( q.invAngle ); r.i32( q.angularScale );
  const CPelBuf org = cs.getOrgBuf( pu.Y() );
  r.u64( refDist( org, pred, bd, false ) );
  r.u64( refDist( org, pred, bd, true ) );
  std::vector<int16_t> v = flat( pred );
  r.u64( fnv1a( v.data(), v.size() ) );
  r.i32( g_curFull );
  if( g_curFull ) for( int16_t s : v ) r.i16( s );
  r.emit( 'P' );
}

void __wrap__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction* ip, ComponentID c, PelBuf& pred, const PredictionUnit& pu )
{
  __real__ZN15IntraPrediction12predIntraAngE11ComponentIDR7AreaBufIsERK14PredictionUnit( ip, c, pred, pu );
  if( g_out && g_inRmd && c == COMPONENT_Y ) emitPred( ip, pred, pu, false );
}

void __wrap__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit( IntraPrediction* ip, ComponentID c, PelBuf& pred, const PredictionUnit& pu )
{
  __real__ZN15IntraPrediction12predIntraMipE11ComponentIDR7AreaBufIsERK14PredictionUnit( ip, c, pred, pu );
  if( g_out && g_inRmd && c == COMPONENT_Y ) emitPred( ip, pred, pu, true );
}

// ---- mode bits -----------------------------------------------------------------------------------------
// 'B': u32 visitId, i32 mip, mode, mrl, u64 fracBits
void __wrap__ZN11CABACWriter20intra_luma_pred_modeERK14PredictionUnit( CABACWriter* cw, const PredictionUnit& pu )
{
  __real__ZN11CABACWriter20intra_luma_pred_modeERK14PredictionUnit( cw, pu );
  if( !g_out || !g_inRmd || !g_is || cw != g_is->m_CABACEstimator ) return;
  Rec r;
  r.u32( g_visitId ); r.i32( pu.cu->mipFlag ); r.i32( pu.intraDir[0] ); r.i32( pu.cu->mipFlag ? 0 : pu.multiRefIdx );
  r.u64( cw->m_BinEncoder.getEstFracBits() );
  r.emit( 'B' );
}

// ---- transforms ----------------------------------------------------------------------------------------
static bool keepTu( int w, int h )
{
  int& n = g_tuShapeCount[w * 256 + h];
  const bool k = n < g_tuFirst || ( n % g_tuStride ) == 0;
  n++;
  return k;
}

// 'S' (MTS pre-selection): i32 w,h,bitDepth,maxCand, nModes, resi[w*h] i16, {i32 mtsIdx, selected, coeff[w*h] i32}*
void __wrap__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi( TrQuant* tq, TransformUnit& tu, const ComponentID& c, const QpParam& qp, std::vector<TrMode>* modes, int maxCand )
{
  g_inRmd = false;
  __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamPSt6vectorISt4pairIibESaISA_EEi( tq, tu, c, qp, modes, maxCand );
  if( !g_out || c != COMPONENT_Y || tu.noResidual || tu.cu->ispMode ) return;
  const CompArea& rect = tu.blocks[c];
  if( !keepTu( rect.width, rect.height ) ) return;
  const int n = rect.width * rect.height;
  Rec r;
  r.i32( rect.width ); r.i32( rect.height ); r.i32( tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) ); r.i32( maxCand ); r.i32( (int) modes->size() );
  putBlock( r, tu.cs->getResiBuf( rect ) );
  for( const auto& m : *modes )
  {
    r.i32( m.first ); r.i32( m.second );
    for( int i = 0; i < n; i++ ) r.i32( tq->m_mtsCoeffs[m.first][i] );
  }
  r.emit( 'S' );
}

// 'Q' (transform + quant): i32 w,h,bitDepth,mtsIdx,lfnstIdx,loadTr,qp,per,rem,absSum,useDQ, f64 lambda, resi[w*h] i16, coeff[w*h] i32, level[w*h] i32
void __wrap__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb( TrQuant* tq, TransformUnit& tu, const ComponentID& c, const QpParam& qp, TCoeff& absSum, const Ctx& ctx, bool loadTr )
{
  g_inRmd = false;
  __real__ZN7TrQuant12transformNxNER13TransformUnitRK11ComponentIDRK7QpParamRiRK3Ctxb( tq, tu, c, qp, absSum, ctx, loadTr );
  if( g_out && c == COMPONENT_Y && !tu.noResidual && !tu.cu->ispMode && tu.cu->lfnstIdx && tu.mtsIdx != MTS_SKIP && tu.cs->slice->getDepQuantEnabledFlag() )
  {
    // 'F' (forward primary transform + forward LFNST + dependent quantisation, CL/TrQuant.cpp:437-560, :1220): i32 w,h,bitDepth,mtsIdx,
    // lfnstIdx,intraMode (PU::getFinalIntraMode, planar for MIP),qp,per,rem,absSum,cbfDeltaBits, f64 lambda, u32 rates[282] (as 'D'),
    // resi[w*h] i16, coeff[w*h] i32 (after LFNST), level[w*h] i32
    const CompArea& rect = tu.blocks[c];
    if( keepTu( rect.width + 3000, rect.height ) )
    {
      const int n = rect.width * rect.height;
      const PredictionUnit& pu = *tu.cs->getPU( rect.pos(), CHANNEL_TYPE_LUMA );
      const int intraMode = PU::isMIP( pu, CHANNEL_TYPE_LUMA ) ? PLANAR_IDX : (int) PU::getFinalIntraMode( pu, CHANNEL_TYPE_LUMA );
      const FracBitsAccess& fb = ctx.getFracBitsAcess();
      Rec d;
      d.i32( rect.width ); d.i32( rect.height ); d.i32( tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) ); d.i32( tu.mtsIdx ); d.i32( tu.cu->lfnstIdx ); d.i32( intraMode );
      d.i32( qp.Qp( false ) ); d.i32( qp.per( false ) ); d.i32( qp.rem( false ) ); d.i32( absSum );
      const BinFracBits cbf = fb.getFracBitsArray( Ctx::QtCbf[COMPONENT_Y]( DeriveCtx::CtxQtCbf( COMPONENT_Y, tu.cbf[COMPONENT_Cb] ) ) );
      d.i32( int32_t( cbf.intBits[1] ) - int32_t( cbf.intBits[0] ) );
      d.f64( tq->m_quant->getLambda() );
      auto put = [&]( const CtxSet& set, int num ) { for( int i = 0; i < num; i++ ) { const BinFracBits b = fb.getFracBitsArray( set( i ) ); d.u32( b.intBits[0] ); d.u32( b.intBits[1] ); } };
      put( Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2 );
      for( int st = 0; st < 3; st++ ) put( Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * st], 12 );
      put( Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21 );
      put( Ctx::LastX[CHANNEL_TYPE_LUMA], 20 ); put( Ctx::LastY[CHANNEL_TYPE_LUMA], 20 );
      putBlock( d, tu.cs->getResiBuf( rect ) );
      const TCoeff* co = loadTr ? tq->m_mtsCoeffs[tu.mtsIdx] : tq->m_tempCoeff;
      for( int i = 0; i < n; i++ ) d.i32( co[i] );
      const CCoeffBuf lv = tu.getCoeffs( c );
      for( int y = 0; y < (int) rect.height; y++ ) for( int x = 0; x < (int) rect.width; x++ ) d.i32( lv.at( x, y ) );
      d.emit( 'F' );
    }
  }
  if( !g_out || c != COMPONENT_Y || tu.noResidual || tu.cu->ispMode || tu.cu->lfnstIdx ) return;
  const CompArea& rect = tu.blocks[c];
  if( !keepTu( rect.width + 1000, rect.height ) ) return;
  const int n = rect.width * rect.height;
  const bool ts = tu.mtsIdx == MTS_SKIP;
  Rec r;
  r.i32( rect.width ); r.i32( rect.height ); r.i32( tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) ); r.i32( tu.mtsIdx ); r.i32( tu.cu->lfnstIdx ); r.i32( loadTr );
  r.i32( qp.Qp( ts ) ); r.i32( qp.per( ts ) ); r.i32( qp.rem( ts ) ); r.i32( absSum );
  r.i32( tu.cs->slice->getDepQuantEnabledFlag() );
  r.f64( tq->m_quant->getLambda() );
  putBlock( r, tu.cs->getResiBuf( rect ) );
  const TCoeff* co = loadTr ? tq->m_mtsCoeffs[tu.mtsIdx] : tq->m_tempCoeff;
  for( int i = 0; i < n; i++ ) r.i32( co[i] );
  const CCoeffBuf lv = tu.getCoeffs( c );
  for( int y = 0; y < (int) rect.height; y++ ) for( int x = 0; x < (int) rect.width; x++ ) r.i32( lv.at( x, y ) );
  r.emit( 'Q' );
  // 'D' (dependent quantisation, CL/DepQuant.cpp:1592): the same call seen from the quantiser -- coefficients in, the context
  // prices its RateEstimator reads, levels out.  i32 w,h,bitDepth,mtsIdx,lfnstIdx,qp,per,rem,absSum,cbfDeltaBits, f64 lambda,
  // u32 sigSbb[2][2], sig[3][12][2], par[21][2], gt1[21][2], gt2[21][2], lastX[20][2], lastY[20][2], resi[w*h] i16, coeff[w*h] i32, level[w*h] i32
  if( tu.cs->slice->getDepQuantEnabledFlag() && !ts )
  {
    const FracBitsAccess& fb = ctx.getFracBitsAcess();
    Rec d;
    d.i32( rect.width ); d.i32( rect.height ); d.i32( tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) ); d.i32( tu.mtsIdx ); d.i32( tu.cu->lfnstIdx );
    d.i32( qp.Qp( false ) ); d.i32( qp.per( false ) ); d.i32( qp.rem( false ) ); d.i32( absSum );
    // RateEstimator::xSetLastCoeffOffset, intra luma without ISP (CL/DepQuant.cpp:531-540)
    const BinFracBits cbf = fb.getFracBitsArray( Ctx::QtCbf[COMPONENT_Y]( DeriveCtx::CtxQtCbf( COMPONENT_Y, tu.cbf[COMPONENT_Cb] ) ) );
    d.i32( int32_t( cbf.intBits[1] ) - int32_t( cbf.intBits[0] ) );
    d.f64( tq->m_quant->getLambda() );
    auto put = [&]( const CtxSet& set, int num ) { for( int i = 0; i < num; i++ ) { const BinFracBits b = fb.getFracBitsArray( set( i ) ); d.u32( b.intBits[0] ); d.u32( b.intBits[1] ); } };
    put( Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2 );
    for( int st = 0; st < 3; st++ ) put( Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * st], 12 );
    put( Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21 );
    put( Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21 );
    put( Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21 );
    put( Ctx::LastX[CHANNEL_TYPE_LUMA], 20 );
    put( Ctx::LastY[CHANNEL_TYPE_LUMA], 20 );
    putBlock( d, tu.cs->getResiBuf( rect ) );
    for( int i = 0; i < n; i++ ) d.i32( co[i] );
    for( int y = 0; y < (int) rect.height; y++ ) for( int x = 0; x < (int) rect.width; x++ ) d.i32( lv.at( x, y ) );
    d.emit( 'D' );
  }
  // 'T' (RDOQ for transform skip, CL/QuantRDOQ.cpp:1243): i32 w,h,bitDepth,qp,per,rem,absSum, f64 lambda,
  // u32 tsSigSbb[3][2], tsSig[3][2], tsPar[1][2], tsGtx[5][2], tsLrg1[4][2], tsSign[6][2], resi[w*h] i16, coeff[w*h] i32, level[w*h] i32
  if( ts && !tu.cu->bdpcmMode )
  {
    const FracBitsAccess& fb = ctx.getFracBitsAcess();
    Rec d;
    d.i32( rect.width ); d.i32( rect.height ); d.i32( tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) );
    d.i32( qp.Qp( true ) ); d.i32( qp.per( true ) ); d.i32( qp.rem( true ) ); d.i32( absSum );
    d.f64( tq->m_quant->getLambda() );
    auto put = [&]( const CtxSet& set, int num ) { for( int i = 0; i < num; i++ ) { const BinFracBits b = fb.getFracBitsArray( set( i ) ); d.u32( b.intBits[0] ); d.u32( b.intBits[1] ); } };
    put( Ctx::TsSigCoeffGroup, 3 ); put( Ctx::TsSigFlag, 3 ); put( Ctx::TsParFlag, 1 ); put( Ctx::TsGtxFlag, 5 ); put( Ctx::TsLrg1Flag, 4 ); put( Ctx::TsResidualSign, 6 );
    putBlock( d, tu.cs->getResiBuf( rect ) );
    for( int i = 0; i < n; i++ ) d.i32( co[i] );
    for( int y = 0; y < (int) rect.height; y++ ) for( int x = 0; x < (int) rect.width; x++ ) d.i32( lv.at( x, y ) );
    d.emit( 'T' );
  }
}

// 'I' (dequant + inverse): i32 w,h,bitDepth,mtsIdx,qp,per,rem, level[w*h] i32, resi[w*h] i16
void __wrap__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam( TrQuant* tq, TransformUnit& tu, const ComponentID& c, PelBuf& resi, const QpParam& qp )
{
  __real__ZN7TrQuant15invTransformNxNER13TransformUnitRK11ComponentIDR7AreaBufIsERK7QpParam( tq, tu, c, resi, qp );
  if( g_out && c == COMPONENT_Y && !tu.cu->ispMode && tu.cu->lfnstIdx && !tu.cu->bdpcmMode && tu.mtsIdx != MTS_SKIP && tu.cs->slice->getDepQuantEnabledFlag() )
  {
    // 'J' (dependent dequantisation + inverse LFNST + inverse primary, CL/TrQuant.cpp:316-435, :561): i32 w,h,bitDepth,mtsIdx,lfnstIdx,
    // intraMode,qp, level[w*h] i32, resi[w*h] i16
    const CompArea& rect = tu.blocks[c];
    if( keepTu( rect.width + 4000, rect.height ) )
    {
      const PredictionUnit& pu = *tu.cs->getPU( rect.pos(), CHANNEL_TYPE_LUMA );
      const int intraMode = PU::isMIP( pu, CHANNEL_TYPE_LUMA ) ? PLANAR_IDX : (int) PU::getFinalIntraMode( pu, CHANNEL_TYPE_LUMA );
      Rec r;
      r.i32( rect.width ); r.i32( rect.height ); r.i32( tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) ); r.i32( tu.mtsIdx ); r.i32( tu.cu->lfnstIdx ); r.i32( intraMode );
      r.i32( qp.Qp( false ) );
      const CCoeffBuf lv = tu.getCoeffs( c );
      for( int y = 0; y < (int) rect.height; y++ ) for( int x = 0; x < (int) rect.width; x++ ) r.i32( lv.at( x, y ) );
      putBlock( r, resi );
      r.emit( 'J' );
    }
  }
  if( !g_out || c != COMPONENT_Y || tu.cu->ispMode || tu.cu->lfnstIdx || tu.cu->bdpcmMode ) return;
  const CompArea& rect = tu.blocks[c];
  if( !keepTu( rect.width + 2000, rect.height ) ) return;
  const bool ts = tu.mtsIdx == MTS_SKIP;
  Rec r;
  r.i32( rect.width ); r.i32( rect.height ); r.i32( tu.cs->sps->getBitDepth( CHANNEL_TYPE_LUMA ) ); r.i32( tu.mtsIdx );
  r.i32( qp.Qp( ts ) ); r.i32( qp.per( ts ) ); r.i32( qp.rem( ts ) );
  const CCoeffBuf lv = tu.getCoeffs( c );
  for( int y = 0; y < (int) rect.height; y++ ) for( int x = 0; x < (int) rect.width; x++ ) r.i32( lv.at( x, y ) );
  putBlock( r, resi );
  r.emit( 'I' );
}

// 'H' (EncCu::updateCtuDataISlice): i32 w,h,result, org[w*h] i16
int __real__ZN5EncCu19updateCtuDataISliceE7AreaBufIKsE( EncCu* cu, const CPelBuf buf );
int __wrap__ZN5EncCu19updateCtuDataISliceE7AreaBufIKsE( EncCu* cu, const CPelBuf buf )
{
  const int res = __real__ZN5EncCu19updateCtuDataISliceE7AreaBufIKsE( cu, buf );
  init();
  if( !g_out ) return res;
  Rec r;
  r.i32( buf.width ); r.i32( buf.height ); r.i32( res );
  for( int y = 0; y < (int) buf.height; y++ ) for( int x = 0; x < (int) buf.width; x++ ) r.i16( buf.at( x, y ) );
  r.emit( 'H' );
  return res;
}

// 'C' (CABACWriter::residual_coding on the bit estimator, EL/CABACWriter.cpp:3773; called from IntraSearch::xEncCoeffQT):
// i32 w,h,mtsIdx,tsAllowed,mtsAllowed,depQuant, u64 fracBits of the call, then the estimator's context states BEFORE the call as
// {u16 state0, u16 state1, u16 rate} for MTSIndex[11], SigCoeffGroup[2], SigFlag[0|2|4][12], ParFlag[21], GtxFlag[2][21] (gt1),
// GtxFlag[0][21] (gt2), LastX[20], LastY[20], TsSigCoeffGroup[3], TsSigFlag[3], TsParFlag[1], TsGtxFlag[5], TsLrg1Flag[4],
// TsResidualSign[6]; level[w*h] i32
void __wrap__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx( CABACWriter* cw, const TransformUnit& tu, ComponentID c, CUCtx* cuCtx )
{
  init();
  const bool rec = g_out && c == COMPONENT_Y && !cw->m_BinEncoder.isEncoding() && !tu.cu->ispMode && !tu.cu->bdpcmMode;
  Rec r;
  uint64_t before = 0;
  if( rec )
  {
    const CompArea& rect = tu.blocks[c];
    r.i32( rect.width ); r.i32( rect.height ); r.i32( tu.mtsIdx ); r.i32( TU::isTSAllowed( tu, c ) ); r.i32( TU::isMTSAllowed( tu, c ) );
    r.i32( tu.cs->slice->getDepQuantEnabledFlag() );
    before = cw->m_BinEncoder.getEstFracBits();
  }
  std::vector<uint16_t> st;
  if( rec )
  {
    const Ctx& ctx = cw->getCtx();
    auto put = [&]( const CtxSet& set, int num ) { for( int i = 0; i < num; i++ ) { const BinProbModel_Std& m = ctx.m_CtxStore_Std[ set( i ) ]; st.push_back( m.m_state[0] ); st.push_back( m.m_state[1] ); st.push_back( m.m_rate ); } };
    put( Ctx::MTSIndex, 11 );
    put( Ctx::SigCoeffGroup[CHANNEL_TYPE_LUMA], 2 );
    for( int k = 0; k < 3; k++ ) put( Ctx::SigFlag[CHANNEL_TYPE_LUMA + 2 * k], 12 );
    put( Ctx::ParFlag[CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[2 + CHANNEL_TYPE_LUMA], 21 ); put( Ctx::GtxFlag[CHANNEL_TYPE_LUMA], 21 );
    put( Ctx::LastX[CHANNEL_TYPE_LUMA], 20 ); put( Ctx::LastY[CHANNEL_TYPE_LUMA], 20 );
    put( Ctx::TsSigCoeffGroup, 3 ); put( Ctx::TsSigFlag, 3 ); put( Ctx::TsParFlag, 1 ); put( Ctx::TsGtxFlag, 5 ); put( Ctx::TsLrg1Flag, 4 ); put( Ctx::TsResidualSign, 6 );
  }
  __real__ZN11CABACWriter15residual_codingERK13TransformUnit11ComponentIDP5CUCtx( cw, tu, c, cuCtx );
  if( !rec ) return;
  const CompArea& rect = tu.blocks[c];
  if( !keepTu( rect.width + 5000 + ( tu.mtsIdx == MTS_SKIP ? 1000 : 0 ), rect.height ) ) return;
  r.u64( cw->m_BinEncoder.getEstFracBits() - before );
  for( uint16_t v : st ) r.put<uint16_t>( v );
  const CCoeffBuf lv = tu.getCoeffs( c );
  for( int y = 0; y < (int) rect.height; y++ ) for( int x = 0; x < (int) rect.width; x++ ) r.i32( lv.at( x, y ) );
  r.emit( 'C' );
}

} // extern "C"

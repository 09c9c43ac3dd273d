// TEST INFRASTRUCTURE: golden-vector generator for vvcb_calc_rd_cost.  Calls the UNMODIFIED reference's RdCost::setLambda and
// RdCost::calcRdCost (CL/RdCost.cpp:63-88) out of oracle/_ref/libvtmref.a and prints lambda, bits, distortion and the cost as the
// bit pattern of the IEEE double.  Built and run by `make -f oracle/Makefile.ref rd_cost`; the output is committed as
// tests/golden/rd_cost.txt.  Never linked into the product library.
#include <cstdio>
#include <cstring>
#include <cmath>
#include "CommonLib/CommonDef.h"
#include "CommonLib/RdCost.h"

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint64_t next() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }

int main()
{
  RdCost rd;
  BitDepths bd; bd.recon[CHANNEL_TYPE_LUMA] = bd.recon[CHANNEL_TYPE_CHROMA] = 10;
  printf( "# lambda(bits of the double) frac_bits distortion cost(bits of the double)\n" );
  for( int i = 0; i < 600; i++ )
  {
    double lambda;
    if( i < 64 ) lambda = 0.57 * pow( 2.0, ( 22 + i % 16 - 12 ) / 3.0 ) * ( 1.0 + ( i / 16 ) * 0.05 );   // the intra lambdas of QP 22..37, scaled as the slice level does
    else         lambda = 0.25 + double( next() % 4000000 ) / 1000.0;
    const uint64_t bits = i % 7 == 0 ? 0 : next() >> ( 24 + next() % 30 );
    const uint64_t dist = i % 11 == 0 ? 0 : next() >> ( 28 + next() % 30 );
    rd.setLambda( lambda, bd );
    rd.saveUnadjustedLambda();   // as EncSlice does after setting the slice lambda: calcRdCost's default argument reads this copy (WCG_EXT)
    const double c = rd.calcRdCost( bits, Distortion( dist ) );
    uint64_t lb, cb; memcpy( &lb, &lambda, 8 ); memcpy( &cb, &c, 8 );
    printf( "%016llx %llu %llu %016llx\n", (unsigned long long) lb, (unsigned long long) bits, (unsigned long long) dist, (unsigned long long) cb );
  }
  return 0;
}

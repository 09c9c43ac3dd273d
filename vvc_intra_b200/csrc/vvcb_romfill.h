// Host-side construction of the device ROM (filter words, per-shape mode parameters, MIP matrices).
#pragma once
#include <string.h>
#include "vvcb_core.cuh"
#include "vvc_rom_tables.h"

namespace vvcb {
inline void fill_rom(Rom& r)
{
  memset(&r, 0, sizeof(r));
  for (int t = 0; t < 2; t++)
    for (int f = 0; f < 32; f++) {
      const int8_t* c = (t ? kIntraGaussFilter : kIntraCubicFilter) + 4 * f;
      r.filt[t][f] = (uint32_t)(uint8_t)c[0] | ((uint32_t)(uint8_t)c[1] << 8) | ((uint32_t)(uint8_t)c[2] << 16) | ((uint32_t)(uint8_t)c[3] << 24);
    }
  for (int lw = 2; lw <= 6; lw++)
    for (int lh = 2; lh <= 6; lh++)
      for (int m = 0; m < VVCB_NUM_LUMA_MODE; m++) r.mode[lw - 2][lh - 2][m] = make_mode_param(1 << lw, 1 << lh, m, 0);
  for (int lw = 2; lw <= 6; lw++)
    for (int lh = 2; lh <= 6; lh++) {
      int n = 0;
      for (int key = 0; key < 6; key++)          // key = is_ver * 3 + PDPC class (0 none, 1 angular, 2 pure hor/ver)
        for (int m = 2; m < VVCB_NUM_LUMA_MODE; m++) {
          const ModeParam& p = r.mode[lw - 2][lh - 2][m];
          const int cls = !p.pdpc ? 0 : (p.angle == 0 ? 2 : 1);
          if (p.is_ver * 3 + cls == key) r.angOrder[lw - 2][lh - 2][n++] = (uint8_t)m;
        }
    }
  memcpy(r.mip4, kMipMatrix4x4, sizeof(r.mip4));
  memcpy(r.mip8, kMipMatrix8x8, sizeof(r.mip8));
  memcpy(r.mip16, kMipMatrix16x16, sizeof(r.mip16));
  memcpy(r.mipOff4, kMipOffset4x4, 18); memcpy(r.mipSh4, kMipShift4x4, 18);
  memcpy(r.mipOff8, kMipOffset8x8, 10); memcpy(r.mipSh8, kMipShift8x8, 10);
  memcpy(r.mipOff16, kMipOffset16x16, 6); memcpy(r.mipSh16, kMipShift16x16, 6);
}

}  // namespace vvcb

// forward transform kernels and quantiser scales (vvcb_tu.cuh's TrRom); T has the TrRom layout
template <class T> inline void fill_tr_rom(T& r)
{
  memset(&r, 0, sizeof(r));
  memcpy(r.dct2, kDct2_4, 32); memcpy(r.dct2 + 16, kDct2_8, 128); memcpy(r.dct2 + 80, kDct2_16, 512);
  memcpy(r.dct2 + 336, kDct2_32, 2048); memcpy(r.dct2 + 1360, kDct2_64, 8192);
  memcpy(r.dct8, kDct8_4, 32); memcpy(r.dct8 + 16, kDct8_8, 128); memcpy(r.dct8 + 80, kDct8_16, 512); memcpy(r.dct8 + 336, kDct8_32, 2048);
  memcpy(r.dst7, kDst7_4, 32); memcpy(r.dst7 + 16, kDst7_8, 128); memcpy(r.dst7 + 80, kDst7_16, 512); memcpy(r.dst7 + 336, kDst7_32, 2048);
  for (int i = 0; i < 12; i++) { r.quantScales[i] = kQuantScales[i]; r.invQuantScales[i] = kInvQuantScales[i]; }
}

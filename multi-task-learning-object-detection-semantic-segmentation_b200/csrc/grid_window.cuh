// Closed-form window of one (regular anchor grid, GT box) pair: the rectangle of grid cells outside of which no anchor
// can reach IoU >= t with the box (derivation in encode_grid.cu).  Host + device: the kernels use it, and
// ssdx_grid_window_probe exposes the same code to the CPU tests, which check it against the full IoU matrix.
#pragma once

#include <math.h>

#include "common.cuh"

struct ssdx_grid_rect {
  int r0, nr, c0, nc;
};

// Window of (grid, GT box) for the IoU bound t, clipped to rows [row_lo, row_hi).  false = empty.
__host__ __device__ inline bool ssdx_grid_window(const ssdx_grid_combo& cb, const float4 gb, const float t,
                                            const int row_lo, const int row_hi, ssdx_grid_rect& o) {
  const float gw = gb.z - gb.x + 1.0f, gh = gb.w - gb.y + 1.0f;
  if (!(gw > 0.0f && gh > 0.0f)) return false;  // empty or NaN box: every IoU is 0 (or NaN) and never wins a '>'
  const float ga = gw * gh;
  float nw = 0.0f, nh = 0.0f;  // overlap extents any pair with iou >= t must reach
  if (ga <= 1e30f) {           // (beyond that the bound arithmetic overflows: no culling)
    const float imin = t * (cb.amin + ga) / (1.0f + t) * 0.99998f;
    const float iwmax = fminf(cb.aw, gw), ihmax = fminf(cb.ah, gh);
    if (imin > iwmax * ihmax * 1.00002f) return false;
    nw = imin / ihmax * 0.99999f;
    nh = imin / iwmax * 0.99999f;
  }
  const float padx = cb.px + 4e-6f * (fabsf(gb.x) + fabsf(gb.z));
  const float pady = cb.py + 4e-6f * (fabsf(gb.y) + fabsf(gb.w));
  float fc0 = ((gb.x + nw - 1.0f - cb.hx - padx) - cb.cx0) * cb.inv_dx;
  float fc1 = ((gb.z - nw + 1.0f + cb.hx + padx) - cb.cx0) * cb.inv_dx;
  float fr0 = ((gb.y + nh - 1.0f - cb.hy - pady) - cb.cy0) * cb.inv_dy;
  float fr1 = ((gb.w - nh + 1.0f + cb.hy + pady) - cb.cy0) * cb.inv_dy;
  fc0 = ceilf(fc0 - (1e-3f + 1e-6f * fabsf(fc0)));
  fc1 = floorf(fc1 + (1e-3f + 1e-6f * fabsf(fc1)));
  fr0 = ceilf(fr0 - (1e-3f + 1e-6f * fabsf(fr0)));
  fr1 = floorf(fr1 + (1e-3f + 1e-6f * fabsf(fr1)));
  // clamp in floating point first: the bounds may be far outside the int range
  const int c0 = (int)fminf(fmaxf(fc0, 0.0f), (float)cb.W);
  const int c1 = (int)fmaxf(fminf(fc1, (float)(cb.W - 1)), -1.0f);
  const int r0 = (int)fminf(fmaxf(fr0, (float)row_lo), (float)row_hi);
  const int r1 = (int)fmaxf(fminf(fr1, (float)(row_hi - 1)), (float)row_lo - 1.0f);
  if (c1 < c0 || r1 < r0) return false;
  o.r0 = r0;
  o.nr = r1 - r0 + 1;
  o.c0 = c0;
  o.nc = c1 - c0 + 1;
  return true;
}


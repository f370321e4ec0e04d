/*
 * sdft_synth.cuh -- K4, the synthesis kernel (c/src/sdft/sdft.h:635-672).
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include "sdft_common.cuh"

namespace sdftb200
{

/* ------------------------------------------------------------------------------------------------
 * K4  synthesis (sdft.h:635-672): one warp per row, lanes stride over bins, shuffle reduction.
 *     latency == 1: y = 2 * sum_k Re(dft[k]) * (-1)^k ; otherwise y = 2 * sum_k Re(dft[k] * tws[k]).
 *     The reference adds bins sequentially; the warp adds them in a fixed tree order instead
 *     (deterministic, differs by rounding only).
 * ---------------------------------------------------------------------------------------------- */
template <typename F>
__device__ __forceinline__ cx<F> load_stream(const cx<F>* p);
template <>
__device__ __forceinline__ cx<double> load_stream<double>(const cx<double>* p)
{
  cx<double> v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.r), "=d"(v.i) : "l"(p));
  return v;
}
template <>
__device__ __forceinline__ cx<float> load_stream<float>(const cx<float>* p)
{
  cx<float> v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.r), "=f"(v.i) : "l"(p));
  return v;
}

__device__ __forceinline__ float4 load_stream4(const float4* p)
{
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

constexpr int kSynthWarps = 8;

template <typename T, typename F, bool UNIT_LATENCY>
__global__ void __launch_bounds__(kSynthWarps * 32) synth_kernel(const cx<F>* __restrict__ dfts,
                                                                 size_t dft_channel_stride,
                                                                 const cx<F>* __restrict__ tws,
                                                                 T* __restrict__ samples, size_t sample_stride,
                                                                 unsigned long long n, unsigned m_all, unsigned roi_first)
{
  /* rows hold `m_all` bins starting at bin `roi_first` (the whole spectrum: roi_first = 0); shifting the
   * weight table keeps the loops below unchanged, the sign of the latency-1 sum follows the absolute bin */
  const unsigned m = m_all;
  tws += roi_first;
  const bool flip = (roi_first & 1u) != 0;
  const unsigned ch = blockIdx.y;
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long warps = (unsigned long long)gridDim.x * kSynthWarps;
  const cx<F>* base = dfts + (size_t)ch * dft_channel_stride;
  T* y = samples + (size_t)ch * sample_stride;
  const bool pairs = (m % 2 == 0) && (reinterpret_cast<uintptr_t>(base) % 16 == 0) && (reinterpret_cast<uintptr_t>(tws) % 16 == 0);
  for (unsigned long long row = (unsigned long long)blockIdx.x * kSynthWarps + (threadIdx.x >> 5); row < n; row += warps)
  {
    const cx<F>* r = base + (size_t)row * m;
    F s0 = (F)0, s1 = (F)0, s2 = (F)0, s3 = (F)0;
    if constexpr (sizeof(F) == sizeof(float))
    {
      /* float rows: two bins (16 bytes) per lane and load when the rows are 16-byte aligned; the pair
       * (even bin, odd bin) carries the signs (+, -) of the latency-1 sum */
      if (pairs)
      {
        const float4* r4 = reinterpret_cast<const float4*>(r);
        const float4* w4 = reinterpret_cast<const float4*>(tws);
        const unsigned mp = m >> 1;
        unsigned q = lane;
        for (; q + 96 < mp; q += 128)
        {
          const float4 a0 = load_stream4(r4 + q), a1 = load_stream4(r4 + q + 32);
          const float4 a2 = load_stream4(r4 + q + 64), a3 = load_stream4(r4 + q + 96);
          if (UNIT_LATENCY)
          {
            s0 += a0.x - a0.z; s1 += a1.x - a1.z; s2 += a2.x - a2.z; s3 += a3.x - a3.z;
          }
          else
          {
            const float4 b0 = w4[q], b1 = w4[q + 32], b2 = w4[q + 64], b3 = w4[q + 96];
            s0 += (a0.x * b0.x - a0.y * b0.y) + (a0.z * b0.z - a0.w * b0.w);
            s1 += (a1.x * b1.x - a1.y * b1.y) + (a1.z * b1.z - a1.w * b1.w);
            s2 += (a2.x * b2.x - a2.y * b2.y) + (a2.z * b2.z - a2.w * b2.w);
            s3 += (a3.x * b3.x - a3.y * b3.y) + (a3.z * b3.z - a3.w * b3.w);
          }
        }
        for (; q < mp; q += 32)
        {
          const float4 a = load_stream4(r4 + q);
          if (UNIT_LATENCY) s0 += a.x - a.z;
          else
          {
            const float4 b = w4[q];
            s0 += (a.x * b.x - a.y * b.y) + (a.z * b.z - a.w * b.w);
          }
        }
        F s = (s0 + s1) + (s2 + s3);
        if (UNIT_LATENCY && flip) s = -s;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) y[row] = (T)(s * (F)2);
        continue;
      }
    }
    unsigned k = lane;
    for (; k + 96 < m; k += 128)
    {
      const cx<F> v0 = load_stream<F>(r + k);
      const cx<F> v1 = load_stream<F>(r + k + 32);
      const cx<F> v2 = load_stream<F>(r + k + 64);
      const cx<F> v3 = load_stream<F>(r + k + 96);
      if (UNIT_LATENCY)
      {
        s0 += v0.r; s1 += v1.r; s2 += v2.r; s3 += v3.r;
      }
      else
      {
        const cx<F> w0 = tws[k], w1 = tws[k + 32], w2 = tws[k + 64], w3 = tws[k + 96];
        s0 += v0.r * w0.r - v0.i * w0.i;
        s1 += v1.r * w1.r - v1.i * w1.i;
        s2 += v2.r * w2.r - v2.i * w2.i;
        s3 += v3.r * w3.r - v3.i * w3.i;
      }
    }
    for (; k < m; k += 32)
    {
      const cx<F> v = load_stream<F>(r + k);
      if (UNIT_LATENCY)
      {
        s0 += v.r;
      }
      else
      {
        const cx<F> w = tws[k];
        s0 += v.r * w.r - v.i * w.i;
      }
    }
    F s = (s0 + s1) + (s2 + s3);
    /* k = lane + 32 i has the parity of the lane: apply (-1)^k once per lane */
    if (UNIT_LATENCY && (((lane & 1) != 0) != flip)) s = -s;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[row] = (T)(s * (F)2);
  }
}


/* ------------------------------------------------------------------------------------------------
 * Window a given (n, m) matrix of UN-windowed rows in the frequency domain: the stand-alone form of
 * sdft_etc_convolve (sdft.h:350-402) that the reference's Python class exposes as SDFT.convolve
 * (python/src/sdft/sdft.py:146-203), mirror cells included (below bin 0 about bin 0, above bin m-1 about
 * bin m-1).  Stateless, one thread per bin, HBM-bound (reads and writes n*m complex values).
 * c0/c1/c2 are the centre / first / second neighbour taps including the caller's scale.
 * ---------------------------------------------------------------------------------------------- */
template <typename F>
__global__ void convolve_kernel(const cx<F>* __restrict__ in, cx<F>* __restrict__ out, unsigned long long n, unsigned m,
                                int window, F c0, F c1, F c2)
{
  const unsigned long long total = n * m;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
  {
    const unsigned long long row = idx / m;
    const int k = (int)(idx - row * m);
    const cx<F>* r = in + row * m;
    auto cell = [&](int j) -> cx<F>
    {
      cx<F> v;
      if (j < 0) { v = r[-j]; v.i = -v.i; }                          // aux[-i] = conj(aux[+i])
      else if (j >= (int)m) { v = r[2 * ((int)m - 1) - j]; v.i = -v.i; }   // aux[(m-1)+i] = conj(aux[(m-1)-i])
      else v = r[j];
      return v;
    };
    const cx<F> c = r[k];
    cx<F> y;
    y.r = c.r * c0; y.i = c.i * c0;
    if (window != 0)
    {
      const cx<F> l1 = cell(k - 1), r1 = cell(k + 1);
      y.r -= (l1.r + r1.r) * c1; y.i -= (l1.i + r1.i) * c1;
      if (window == 3)
      {
        const cx<F> l2 = cell(k - 2), r2 = cell(k + 2);
        y.r += (l2.r + r2.r) * c2; y.i += (l2.i + r2.i) * c2;
      }
    }
    out[idx] = y;
  }
}

/* scattered row pointers (sdft_sdft_nd / sdft_isdft_nd, sdft.h:622-628, :681-687): tile (rows, m) <-> rows[r][0..m),
 * rows with a null pointer are skipped (they live in host memory and take the staging path) */
template <typename F>
__global__ void scatter_rows_kernel(const cx<F>* __restrict__ tile, cx<F>* const* __restrict__ rows, unsigned long long nrows, unsigned m)
{
  const unsigned long long total = nrows * m, stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
  {
    const unsigned long long r = idx / m;
    cx<F>* dst = rows[r];
    if (dst) dst[idx - r * m] = tile[idx];
  }
}
template <typename F>
__global__ void gather_rows_kernel(cx<F>* __restrict__ tile, const cx<F>* const* __restrict__ rows, unsigned long long nrows, unsigned m)
{
  const unsigned long long total = nrows * m, stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
  {
    const unsigned long long r = idx / m;
    const cx<F>* src = rows[r];
    if (src) tile[idx] = src[idx - r * m];
  }
}

}  // namespace sdftb200

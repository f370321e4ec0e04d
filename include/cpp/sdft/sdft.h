/*
 * sdft/sdft.h (C++) -- drop-in replacement for the C++ template of jurihock/sdft
 * (cpp/src/sdft/sdft.h:29-352): sdft::SDFT<T, F> with the same constructor, reset/size/window/latency
 * and the three sdft / three isdft overloads, but every method forwards to libsdft_b200.so
 * (include/sdft_b200.h) and runs on a B200.  Put include/cpp on the include path instead of the
 * reference's cpp/src and link -lsdft_b200.
 *
 * T and F may each be float or double.  long double is rejected at compile time: there is no 80-bit
 * arithmetic on the GPU.  (The reference itself only compiles F = double / long double under
 * libstdc++; F = float works here because the tables are produced by the C-ABI library.)
 * Construction throws std::runtime_error when no CUDA device is usable -- there is no CPU fallback.
 */
#pragma once

#include <complex>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "../../sdft_b200.h"

namespace sdft
{
  /** Supported SDFT analysis window types (cpp/src/sdft/sdft.h:34-40). */
  enum class Window
  {
    Boxcar,
    Hann,
    Hamming,
    Blackman
  };

  namespace detail
  {
    template <typename T, typename F> struct abi;

#define SDFT_B200_CPP_ABI(T, F, SFX, FDX)                                                                    \
    template <> struct abi<T, F>                                                                             \
    {                                                                                                        \
      typedef FDX fdx;                                                                                       \
      static sdft_b200_plan_t* alloc(size_t m, int w, double l) { return sdft_b200_##SFX##_alloc_custom(m, w, l); } \
      static void free(sdft_b200_plan_t* p) { sdft_b200_##SFX##_free(p); }                                   \
      static void reset(sdft_b200_plan_t* p) { sdft_b200_##SFX##_reset(p); }                                 \
      static void sdft(sdft_b200_plan_t* p, T x, fdx* d) { sdft_b200_##SFX##_sdft(p, x, d); }                \
      static void sdft_n(sdft_b200_plan_t* p, size_t n, const T* x, fdx* d) { sdft_b200_##SFX##_sdft_n(p, n, x, d); } \
      static void sdft_nd(sdft_b200_plan_t* p, size_t n, const T* x, fdx** d) { sdft_b200_##SFX##_sdft_nd(p, n, x, d); } \
      static T isdft(sdft_b200_plan_t* p, const fdx* d) { return sdft_b200_##SFX##_isdft(p, d); }            \
      static void isdft_n(sdft_b200_plan_t* p, size_t n, const fdx* d, T* y) { sdft_b200_##SFX##_isdft_n(p, n, d, y); } \
      static void isdft_nd(sdft_b200_plan_t* p, size_t n, const fdx** d, T* y) { sdft_b200_##SFX##_isdft_nd(p, n, d, y); } \
    };

    SDFT_B200_CPP_ABI(float, float, f32f32, sdft_b200_cf32_t)
    SDFT_B200_CPP_ABI(float, double, f32f64, sdft_b200_cf64_t)
    SDFT_B200_CPP_ABI(double, float, f64f32, sdft_b200_cf32_t)
    SDFT_B200_CPP_ABI(double, double, f64f64, sdft_b200_cf64_t)
#undef SDFT_B200_CPP_ABI
  }

  /**
   * Sliding Discrete Fourier Transform (SDFT) on a B200.
   * @tparam T Time domain data type: float (default) or double.
   * @tparam F Frequency domain data type: float or double (default and recommended).
   **/
  template <typename T = float, typename F = double>
  class SDFT
  {
    static_assert((std::is_same<T, float>::value || std::is_same<T, double>::value) &&
                  (std::is_same<F, float>::value || std::is_same<F, double>::value),
                  "sdft_b200: T and F must be float or double (long double is not supported on the GPU)");

    typedef detail::abi<T, F> abi;
    typedef typename abi::fdx fdx;

  public:

    /** Creates a new SDFT plan (cpp/src/sdft/sdft.h:61). */
    SDFT(const size_t dftsize, const Window window = Window::Hann, const double latency = 1) :
      plan(abi::alloc(dftsize, static_cast<int>(window), latency)),
      dftsize(dftsize),
      windowtype(window),
      latencyfactor(latency)
    {
      if (plan == nullptr)
      {
        throw std::runtime_error(std::string("sdft_b200: ") + sdft_b200_last_error_string(nullptr));
      }
    }

    ~SDFT() { abi::free(plan); }

    SDFT(const SDFT&) = delete;
    SDFT& operator=(const SDFT&) = delete;

    /** Resets this SDFT plan instance to its initial state (cpp/src/sdft/sdft.h:97). */
    void reset() { abi::reset(plan); }

    /** Returns the assigned number of DFT bins. */
    size_t size() const { return dftsize; }

    /** Returns the assigned analysis window type. */
    Window window() const { return windowtype; }

    /** Returns the assigned synthesis latency factor. */
    double latency() const { return latencyfactor; }

    /** Estimates the DFT vector for the given sample (cpp/src/sdft/sdft.h:135). */
    void sdft(const T sample, std::complex<F>* const dft)
    {
      abi::sdft(plan, sample, reinterpret_cast<fdx*>(dft));
    }

    /** Estimates the DFT matrix (nsamples, dftsize) for the given sample array (cpp/src/sdft/sdft.h:179). */
    void sdft(const size_t nsamples, const T* samples, std::complex<F>* const dfts)
    {
      abi::sdft_n(plan, nsamples, samples, reinterpret_cast<fdx*>(dfts));
    }

    /** Same, into nsamples separately allocated DFT vectors (cpp/src/sdft/sdft.h:193). */
    void sdft(const size_t nsamples, const T* samples, std::complex<F>** const dfts)
    {
      abi::sdft_nd(plan, nsamples, samples, reinterpret_cast<fdx**>(dfts));
    }

    /** Synthesizes a single sample from the given DFT vector (cpp/src/sdft/sdft.h:205). */
    T isdft(const std::complex<F>* dft)
    {
      return abi::isdft(plan, reinterpret_cast<const fdx*>(dft));
    }

    /** Synthesizes the sample array from the given DFT matrix (cpp/src/sdft/sdft.h:235). */
    void isdft(const size_t nsamples, const std::complex<F>* dfts, T* const samples)
    {
      abi::isdft_n(plan, nsamples, reinterpret_cast<const fdx*>(dfts), samples);
    }

    /** Same, from nsamples separately allocated DFT vectors (cpp/src/sdft/sdft.h:249). */
    void isdft(const size_t nsamples, const std::complex<F>** dfts, T* const samples)
    {
      abi::isdft_nd(plan, nsamples, reinterpret_cast<const fdx**>(dfts), samples);
    }

    /** The underlying C-ABI plan, for the extensions of sdft_b200.h (streams, batches, round trip). */
    sdft_b200_plan_t* handle() const { return plan; }

  private:

    sdft_b200_plan_t* const plan;
    const size_t dftsize;
    const Window windowtype;
    const double latencyfactor;

  };
}

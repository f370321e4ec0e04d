/*
 * sdft_kernels.cuh -- sm_100a kernels of the sliding-DFT hot path.
 *
 * Reference being replaced: the per-sample loops of c/src/sdft/sdft.h:562-598 (analysis),
 * :350-402 (window convolution) and :635-672 (synthesis).  See DESIGN.md for the scan formulation.
 *
 * Vocabulary
 *   bin k            0..m-1, the reference's dft index
 *   cell e = k + 2   "extended" index 0..m+3; cells 0,1,m+2,m+3 are the mirror cells of
 *                    sdft.h:589-595.  They are carried as ordinary bins with conjugated twiddles:
 *                    (a+bi)(c+di) and its conjugate round identically, so a mirror cell evolves as the
 *                    exact conjugate of its source bin and no mirroring step exists on the device.
 *   phase P[c][e]    the reference's "fiddle": tw^c by sequential multiplication, restarted every
 *                    2m samples (sdft.h:566-576)
 *   chunk            a run of <= L consecutive samples that never crosses a multiple of L inside the
 *                    2m period nor the period end; every chunk but the first of a call therefore
 *                    starts at a cursor that is a multiple of L and reads its phase from the F0 table
 *
 * Arithmetic policy
 *   float   reproduces every rounding point of the reference (un-fused, packed FMUL2/FADD2): phases
 *           are bit-exact and rows are bit-exact within a chunk; only the summation order across
 *           chunks differs.
 *   double  MODE_MODULATED: the reference's scheme with explicit FMAs in a fixed pattern.
 *           MODE_FAST (default): the chunk total is a Horner sum of tw^i * delta_i scaled by the phase
 *           at the chunk start, and the replay runs the demodulated recurrence
 *           aux <- (aux + delta) * conj(tw) anchored at carry * conj(P_start) at every chunk start,
 *           with the window weight folded into the deltas.  13 instead of 22 FP64 instructions per
 *           bin-update (Hann); the B200 is power-capped on this path, so fewer FP64 operations is
 *           more bandwidth.  Mathematically identical, differs by rounding (~1e-12 of full scale,
 *           gate is 1e-9); the phase still restarts exactly every 2m samples.
 */
#pragma once

#include "sdft_common.cuh"
#include "sdft_arith.cuh"
#include "sdft_schedule.cuh"
#include "sdft_lane.cuh"
#include "sdft_scan.cuh"
#include "sdft_synth.cuh"
#include "sdft_peak.cuh"

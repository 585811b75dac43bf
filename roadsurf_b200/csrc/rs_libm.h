// rs_libm.h -- exp and log evaluated exactly as the host's libm does (glibc >= 2.28, x86-64 FMA code
// path), for host and device code.
//
// The reference is Fortran linked against the system libm.  A last-bit difference in exp or log does
// not matter for the temperatures, but it decides the sign of a rounding residual in the storage
// terms and with it a threshold-induced state flip.  Evaluating both functions
// with libm's own algorithm makes the CUDA path bit-identical to the CPU restatement wherever only
// + - * / sqrt exp log are involved.
//
// PROVENANCE (third-party, not from the RoadSurf reference): the algorithm and the table / coefficient
// data are those of the GNU C Library's double-precision exp and log, sysdeps/ieee754/dbl-64/e_exp.c,
// e_log.c, e_exp_data.c, e_log_data.c -- contributed to glibc from ARM's optimized-routines (Szabolcs
// Nagy; MIT OR Apache-2.0 WITH LLVM-exception upstream, LGPL-2.1-or-later as distributed in glibc).
// rs_libm_tables.h holds numbers read out of the installed libm.so.6, this file re-expresses the
// evaluation scheme; neither contains glibc source text.  Table-driven, 128 entries each.  The ORDER OF OPERATIONS and the placement of the
// fused multiply-adds below are those of libm's __exp_fma / __log_fma as disassembled from the
// libm.so.6 the tables come from (scripts/gen_libm_tables.py); tests/test_libm_match.py compares the
// host build of this header with libm on 2e7 arguments.  Arguments outside the fast path (|x| >= 512
// or NaN for exp; x <= 0, subnormal, inf, NaN for log) take the caller-supplied fallback.
#pragma once

#include <cstdint>
#include <cstring>

#include "rs_libm_tables.h"

#if defined(__CUDACC__)
#define RS_LIBM_HD __host__ __device__ __forceinline__
#else
#define RS_LIBM_HD inline
#endif

namespace rslibm
{
// the 128-entry tables are indexed per lane: global memory through the read-only path (one 16-byte
// load per call); the scalar coefficients are literals
#if defined(__CUDACC__)
__device__ const unsigned long long __align__(16) d_exp_tab[256] = RS_LIBM_EXP_TAB;  // {tail, scale bits}
__device__ const double __align__(16) d_log_tab[256] = RS_LIBM_LOG_TAB;              // {invc, logc}
#endif
static const unsigned long long h_exp_tab[256] = RS_LIBM_EXP_TAB;
static const double h_log_tab[256] = RS_LIBM_LOG_TAB;
// Shared-memory copies of the two tables for kernels that ask for them (template argument SMT = true on
// the functions below): the block copies both tables (4 KB) first (stage_tables, all threads of the block,
// before any of them exits) and the lookups are LDS.128, ~30 cycles instead of an L1 (~40) or L2 (~250)
// access -- which a warp that is alone on its scheduler waits out in full.  Kernels that do not reference
// the arrays allocate nothing.
#if defined(__CUDACC__)
__shared__ ulonglong2 s_exp_tab[128];
__shared__ double2 s_log_tab[128];
__device__ __forceinline__ void stage_tables()
{
  for (unsigned k = threadIdx.x; k < 128u; k += blockDim.x)
  {
    s_exp_tab[k] = __ldg(reinterpret_cast<const ulonglong2*>(d_exp_tab) + k);
    s_log_tab[k] = __ldg(reinterpret_cast<const double2*>(d_log_tab) + k);
  }
  __syncthreads();
}
#endif
// scalar coefficients: on the device operands from the constant bank (as literals each costs two
// move instructions per use and the functions are inlined at several sites), on the host literals
#if defined(__CUDACC__)
__constant__ double c_exp_k[8] = {RS_LIBM_EXP_H0, RS_LIBM_EXP_H1, RS_LIBM_EXP_H2, RS_LIBM_EXP_H3,
                                  RS_LIBM_EXP_H4, RS_LIBM_EXP_H5, RS_LIBM_EXP_H6, RS_LIBM_EXP_H7};
__constant__ double c_log_k[18] = {RS_LIBM_LOG_H0,  RS_LIBM_LOG_H1,  RS_LIBM_LOG_H2,  RS_LIBM_LOG_H3,  RS_LIBM_LOG_H4,
                                   RS_LIBM_LOG_H5,  RS_LIBM_LOG_H6,  RS_LIBM_LOG_H7,  RS_LIBM_LOG_H8,  RS_LIBM_LOG_H9,
                                   RS_LIBM_LOG_H10, RS_LIBM_LOG_H11, RS_LIBM_LOG_H12, RS_LIBM_LOG_H13, RS_LIBM_LOG_H14,
                                   RS_LIBM_LOG_H15, RS_LIBM_LOG_H16, RS_LIBM_LOG_H17};
#endif
#if defined(__CUDA_ARCH__)
#define RS_EK(i) rslibm::c_exp_k[i]
#define RS_LK(i) rslibm::c_log_k[i]
#else
#define RS_EK(i) RS_LIBM_EXP_H##i
#define RS_LK(i) RS_LIBM_LOG_H##i
#endif

RS_LIBM_HD double fma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
RS_LIBM_HD double mul_(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
RS_LIBM_HD double add_(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
RS_LIBM_HD uint64_t asu(double x)
{
#if defined(__CUDA_ARCH__)
  return static_cast<uint64_t>(__double_as_longlong(x));
#else
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return u;
#endif
}
RS_LIBM_HD double asd(uint64_t u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(static_cast<long long>(u));
#else
  double x;
  std::memcpy(&x, &u, 8);
  return x;
#endif
}
template <bool SMT = false>
RS_LIBM_HD void exp_entry(uint32_t i, double& tail, uint64_t& sbits)
{
#if defined(__CUDA_ARCH__)
  ulonglong2 e;
  if constexpr (SMT)
    e = s_exp_tab[i];
  else
    e = __ldg(reinterpret_cast<const ulonglong2*>(d_exp_tab) + i);
  tail = asd(e.x);
  sbits = e.y;
#else
  tail = asd(h_exp_tab[2 * i]);
  sbits = h_exp_tab[2 * i + 1];
#endif
}
template <bool SMT = false>
RS_LIBM_HD void log_entry(uint32_t i, double& invc, double& logc)
{
#if defined(__CUDA_ARCH__)
  double2 e;
  if constexpr (SMT)
    e = s_log_tab[i];
  else
    e = __ldg(reinterpret_cast<const double2*>(d_log_tab) + i);
  invc = e.x;
  logc = e.y;
#else
  invc = h_log_tab[2 * i];
  logc = h_log_tab[2 * i + 1];
#endif
}

// exp(x).  `ok` is false when x is outside the fast path; the caller then uses its fallback.
template <bool SMT = false>
RS_LIBM_HD double exp_fast(double x, bool& ok)
{
  const uint32_t abstop = static_cast<uint32_t>(asu(x) >> 52) & 0x7ff;
  ok = true;
  if (abstop - 0x3c9u > 0x3eu)  // |x| < 2^-54 or |x| >= 512 (or NaN / inf)
  {
    if (abstop < 0x3c9u) return add_(1.0, x);
    ok = false;
    return x;
  }
  const double InvLn2N = RS_EK(0), Shift = RS_EK(1);
  const double NegLn2hiN = RS_EK(2), NegLn2loN = RS_EK(3);
  const double C2 = RS_EK(4), C3 = RS_EK(5), C4 = RS_EK(6),
               C5 = RS_EK(7);
  double kd = fma_(x, InvLn2N, Shift);
  const uint64_t ki = asu(kd);
  kd = add_(kd, -Shift);
  const double r = fma_(kd, NegLn2loN, fma_(kd, NegLn2hiN, x));
  double tail;
  uint64_t sbits;
  exp_entry<SMT>(static_cast<uint32_t>(ki & 127u), tail, sbits);
  sbits += ki << 45;
  const double p23 = fma_(r, C3, C2);
  const double tr = add_(r, tail);
  const double r2 = mul_(r, r);
  const double p45 = fma_(r, C5, C4);
  const double t1 = fma_(p23, r2, tr);
  const double r4 = mul_(r2, r2);
  const double tmp = fma_(r4, p45, t1);
  const double scale = asd(sbits);
  return fma_(scale, tmp, scale);
}

// exp_fast without branches (same operations, same results): every lane evaluates the main path, the
// |x| < 2^-54 case is a select and arguments outside the fast path only clear `ok`.  For call sites that
// are to stay inside one basic block.
template <bool SMT = false>
RS_LIBM_HD double exp_fast_flat(double x, bool& ok)
{
  const uint32_t abstop = static_cast<uint32_t>(asu(x) >> 52) & 0x7ff;
  const bool tiny = abstop < 0x3c9u;
  ok = tiny || (abstop - 0x3c9u <= 0x3eu);
  const double InvLn2N = RS_EK(0), Shift = RS_EK(1);
  const double NegLn2hiN = RS_EK(2), NegLn2loN = RS_EK(3);
  const double C2 = RS_EK(4), C3 = RS_EK(5), C4 = RS_EK(6), C5 = RS_EK(7);
  double kd = fma_(x, InvLn2N, Shift);
  const uint64_t ki = asu(kd);
  kd = add_(kd, -Shift);
  const double r = fma_(kd, NegLn2loN, fma_(kd, NegLn2hiN, x));
  double tail;
  uint64_t sbits;
  exp_entry<SMT>(static_cast<uint32_t>(ki & 127u), tail, sbits);
  sbits += ki << 45;
  const double p23 = fma_(r, C3, C2);
  const double tr = add_(r, tail);
  const double r2 = mul_(r, r);
  const double p45 = fma_(r, C5, C4);
  const double t1 = fma_(p23, r2, tr);
  const double r4 = mul_(r2, r2);
  const double tmp = fma_(r4, p45, t1);
  const double scale = asd(sbits);
  const double y = fma_(scale, tmp, scale);
  return tiny ? add_(1.0, x) : y;
}

// log(x).  `ok` is false for x <= 0, subnormal, inf, NaN.
template <bool SMT = false>
RS_LIBM_HD double log_fast(double x, bool& ok)
{
  const uint64_t ix = asu(x);
  ok = true;
  if (ix - 0x3fee000000000000ull < 0x0003090000000000ull)  // 1 - 2^-4 <= x < 1 + 0x1.09p-4
  {
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = add_(x, -1.0);
    const double B0 = RS_LK(7);
    const double b12 = fma_(r, RS_LK(9), RS_LK(8));     // B1 + r B2
    const double b45 = fma_(r, RS_LK(12), RS_LK(11));   // B4 + r B5
    const double r2 = mul_(r, r);
    const double b78 = fma_(r, RS_LK(15), RS_LK(14));   // B7 + r B8
    const double b123 = fma_(r2, RS_LK(10), b12);
    const double b456 = fma_(r2, RS_LK(13), b45);
    const double r3 = mul_(r, r2);
    double q = fma_(r2, RS_LK(16), b78);   // B7 + r B8 + r2 B9
    q = fma_(r3, RS_LK(17), q);            // + r3 B10
    q = fma_(q, r3, b456);
    q = fma_(q, r3, b123);
    // hi + lo = r - r*r/2 in extra precision
    const double t = fma_(r, 0x1p27, r);
    const double rhi = fma_(-0x1p27, r, t);
    const double rhi2 = mul_(rhi, rhi);
    const double rlo = add_(r, -rhi);
    const double hi = fma_(rhi2, B0, r);
    const double lo = fma_(rhi2, B0, add_(r, -hi));
    const double lo2 = fma_(mul_(B0, rlo), add_(r, rhi), lo);
    const double y = fma_(q, r3, lo2);
    return add_(hi, y);
  }
  const uint32_t top = static_cast<uint32_t>(ix >> 48);
  if (top - 0x0010u > 0x7fdfu)
  {
    ok = false;
    return x;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ull;
  const uint32_t i = static_cast<uint32_t>(tmp >> 45) & 127u;
  const int k = static_cast<int>(static_cast<int64_t>(tmp) >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
  double invc, logc;
  log_entry<SMT>(i, invc, logc);
  const double z = asd(iz);
  const double kd = static_cast<double>(k);
  const double Ln2hi = RS_LK(0), Ln2lo = RS_LK(1);
  const double A0 = RS_LK(2), A1 = RS_LK(3), A2 = RS_LK(4),
               A3 = RS_LK(5), A4 = RS_LK(6);
  const double w = fma_(kd, Ln2hi, logc);
  const double r = fma_(z, invc, -1.0);
  const double a12 = fma_(r, A2, A1);
  const double hi = add_(r, w);
  const double r2 = mul_(r, r);
  const double lo = fma_(kd, Ln2lo, add_(add_(w, -hi), r));
  const double r3 = mul_(r, r2);
  const double a34 = fma_(r, A4, A3);
  const double lo2 = fma_(r2, A0, lo);
  const double p = fma_(a34, r2, a12);
  const double y = fma_(r3, p, lo2);
  return add_(y, hi);
}
}  // namespace rslibm

/*
 * mg_sincos.h — deterministic double-precision sin/cos.
 *
 * The physics needs one (cos a, sin a) per body per sub-step (Chipmunk's
 * cpvforangle in cpBodyUpdatePosition).  libm on the host and CUDA's device
 * libm agree only to 1-2 ulp, which would make bit-exact CPU/GPU parity
 * impossible.  This routine uses only IEEE +,*,fma and round-to-nearest
 * integer conversion, in a fixed order, so the same source gives the same bits
 * under gcc (-ffp-contract=off) and nvcc (any -fmad setting: every multiply /
 * add below goes through a non-contractable intrinsic on the device).
 *
 * Accuracy: Cody-Waite 3-term reduction by pi/2 (exact for |a| < 2^17 * pi/2,
 * far beyond any heading the arena can accumulate) + the classic degree-13/14
 * minimax kernels; max error < 1 ulp-ish (2e-16 absolute), checked in
 * tests/test_sincos.py against libm.
 */
#ifndef MG_SINCOS_H
#define MG_SINCOS_H

#include <math.h>

#if defined(__CUDA_ARCH__)
#define MG_SC_MUL(a, b) __dmul_rn((a), (b))
#define MG_SC_ADD(a, b) __dadd_rn((a), (b))
#define MG_SC_FMA(a, b, c) __fma_rn((a), (b), (c))
#define MG_SC_RINT(a) rint(a)
#define MG_SC_FN __host__ __device__ static __forceinline__
#elif defined(__CUDACC__)
#define MG_SC_MUL(a, b) ((a) * (b))
#define MG_SC_ADD(a, b) ((a) + (b))
#define MG_SC_FMA(a, b, c) fma((a), (b), (c))
#define MG_SC_RINT(a) rint(a)
#define MG_SC_FN __host__ __device__ static __forceinline__
#else
#define MG_SC_MUL(a, b) ((a) * (b))
#define MG_SC_ADD(a, b) ((a) + (b))
#define MG_SC_FMA(a, b, c) fma((a), (b), (c))
#define MG_SC_RINT(a) rint(a)
#define MG_SC_FN static inline
#endif

MG_SC_FN void mg_det_sincos(double a, double* s_out, double* c_out) {
  const double two_over_pi = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
  const double pio2_2 = 6.07710050630396597660e-11;  /* next 33 bits */
  const double pio2_3 = 2.02226624871116645580e-21;  /* next 33 bits */
  const double pio2_3t = 8.47842766036889956997e-32; /* tail */
  double kf = MG_SC_RINT(MG_SC_MUL(a, two_over_pi));
  /* each fma is exact-product + one rounding; the first two are exact
   * subtractions for |k| < 2^17 */
  double r = MG_SC_FMA(-kf, pio2_1, a);
  r = MG_SC_FMA(-kf, pio2_2, r);
  r = MG_SC_FMA(-kf, pio2_3, r);
  r = MG_SC_FMA(-kf, pio2_3t, r);
  double z = MG_SC_MUL(r, r);
  /* sin kernel */
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double ps = MG_SC_FMA(z, S6, S5);
  ps = MG_SC_FMA(z, ps, S4);
  ps = MG_SC_FMA(z, ps, S3);
  ps = MG_SC_FMA(z, ps, S2);
  ps = MG_SC_FMA(z, ps, S1);
  double sr = MG_SC_FMA(MG_SC_MUL(z, r), ps, r);
  /* cos kernel */
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double pc = MG_SC_FMA(z, C6, C5);
  pc = MG_SC_FMA(z, pc, C4);
  pc = MG_SC_FMA(z, pc, C3);
  pc = MG_SC_FMA(z, pc, C2);
  pc = MG_SC_FMA(z, pc, C1);
  double cr = MG_SC_FMA(MG_SC_MUL(z, z), pc, MG_SC_FMA(z, -0.5, 1.0));
  long long k = (long long)kf;
  switch ((int)(k & 3)) {
    case 0: *s_out = sr; *c_out = cr; break;
    case 1: *s_out = cr; *c_out = -sr; break;
    case 2: *s_out = -sr; *c_out = -cr; break;
    default: *s_out = -cr; *c_out = sr; break;
  }
}

#endif /* MG_SINCOS_H */

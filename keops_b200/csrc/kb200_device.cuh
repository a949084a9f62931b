// kb200_device.cuh -- device-side prelude prepended to every generated sm_100a map-reduce kernel.
//
// The generator (keops_b200/codegen/emit.py) defines, before this file:
//   KB_T        threads per CTA                     KB_R     output rows per thread
//   KB_F64      0: fp32 formula, 1: fp64 formula    KB_FAST  1: use_fast_math lowering
//   KB_TJ       streamed rows per shared-memory tile, KB_STRIDE  reals per staged row
//
// Each thread owns KB_R output rows ("lanes").  In fp32 the lanes are held two by two in 64-bit
// register pairs and all per-pair arithmetic is issued as packed FADD2 / FMUL2 / FFMA2 (PTX
// add/sub/mul/fma .f32x2, new on sm_100): one issue slot feeds two (i,j) pairs, streamed
// (j-side) operands enter through the instruction's scalar-broadcast operand form, so the
// staged y_j tile needs no duplication.  In fp64 a "pack" is a single double.
//
// Nothing here is compiled for, or dispatched to, any other architecture.
#pragma once

#if KB_F64
typedef double real;
typedef double pk_t;
#define KB_LPP 1
#else
typedef float real;
typedef float2 pk_t;
#define KB_LPP 2
#endif
#define KB_NP (KB_R / KB_LPP)

#define KB_DEV __device__ __forceinline__
#define KB_UNROLL_ALL _Pragma("unroll")

union V {
  pk_t p[KB_NP];
  real l[KB_R];
};

// ---------------------------------------------------------------------------------------------
// pack-level arithmetic
// ---------------------------------------------------------------------------------------------
#if KB_F64
KB_DEV pk_t pk_bc(real s) { return s; }
KB_DEV pk_t pk_add(pk_t a, pk_t b) { return a + b; }
KB_DEV pk_t pk_sub(pk_t a, pk_t b) { return a - b; }
KB_DEV pk_t pk_mul(pk_t a, pk_t b) { return a * b; }
KB_DEV pk_t pk_fma(pk_t a, pk_t b, pk_t c) { return fma(a, b, c); }
#else
KB_DEV unsigned long long pk_u(pk_t a) { return *reinterpret_cast<unsigned long long*>(&a); }
KB_DEV pk_t pk_f(unsigned long long u) { return *reinterpret_cast<pk_t*>(&u); }
KB_DEV pk_t pk_bc(real s) { return make_float2(s, s); }
#if KB_FAST
KB_DEV pk_t pk_add(pk_t a, pk_t b) { unsigned long long d; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b))); return pk_f(d); }
KB_DEV pk_t pk_sub(pk_t a, pk_t b) { unsigned long long d; asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b))); return pk_f(d); }
KB_DEV pk_t pk_mul(pk_t a, pk_t b) { unsigned long long d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b))); return pk_f(d); }
KB_DEV pk_t pk_fma(pk_t a, pk_t b, pk_t c) { unsigned long long d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b)), "l"(pk_u(c))); return pk_f(d); }
#else
KB_DEV pk_t pk_add(pk_t a, pk_t b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b))); return pk_f(d); }
KB_DEV pk_t pk_sub(pk_t a, pk_t b) { unsigned long long d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b))); return pk_f(d); }
KB_DEV pk_t pk_mul(pk_t a, pk_t b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b))); return pk_f(d); }
KB_DEV pk_t pk_fma(pk_t a, pk_t b, pk_t c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk_u(a)), "l"(pk_u(b)), "l"(pk_u(c))); return pk_f(d); }
#endif
#endif

KB_DEV pk_t kb_pk(const V& v, int p) { return v.p[p]; }
KB_DEV pk_t kb_pk(real s, int) { return pk_bc(s); }
KB_DEV real kb_ln(const V& v, int q) { return v.l[q]; }
KB_DEV real kb_ln(real s, int) { return s; }

KB_DEV V kb_bcast_v(real s) { V r; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) r.l[q] = s; return r; }

template <class A, class B> KB_DEV V kb_add_v(const A& a, const B& b) { V r; KB_UNROLL_ALL for (int p = 0; p < KB_NP; p++) r.p[p] = pk_add(kb_pk(a, p), kb_pk(b, p)); return r; }
template <class A, class B> KB_DEV V kb_sub_v(const A& a, const B& b) { V r; KB_UNROLL_ALL for (int p = 0; p < KB_NP; p++) r.p[p] = pk_sub(kb_pk(a, p), kb_pk(b, p)); return r; }
template <class A, class B> KB_DEV V kb_mul_v(const A& a, const B& b) { V r; KB_UNROLL_ALL for (int p = 0; p < KB_NP; p++) r.p[p] = pk_mul(kb_pk(a, p), kb_pk(b, p)); return r; }
template <class A, class B, class C> KB_DEV V kb_fma_v(const A& a, const B& b, const C& c) { V r; KB_UNROLL_ALL for (int p = 0; p < KB_NP; p++) r.p[p] = pk_fma(kb_pk(a, p), kb_pk(b, p), kb_pk(c, p)); return r; }
KB_DEV V kb_neg_v(const V& a) { return kb_sub_v((real)0, a); }

KB_DEV real kb_add(real a, real b) { return a + b; }
KB_DEV real kb_sub(real a, real b) { return a - b; }
KB_DEV real kb_mul(real a, real b) { return a * b; }
KB_DEV real kb_neg(real a) { return -a; }
#if KB_F64
KB_DEV real kb_fma(real a, real b, real c) { return fma(a, b, c); }
#else
KB_DEV real kb_fma(real a, real b, real c) { return fmaf(a, b, c); }
#endif

// ---------------------------------------------------------------------------------------------
// scalar math (fast = the reference's default -use_fast_math lowering; else IEEE / libdevice)
// ---------------------------------------------------------------------------------------------
#if KB_F64
KB_DEV real kb_rcp(real x) { return 1.0 / x; }
KB_DEV real kb_div(real a, real b) { return a / b; }
KB_DEV real kb_exp(real x) { return exp(x); }
KB_DEV real kb_ex2(real x) { return exp2(x); }
KB_DEV real kb_log(real x) { return log(x); }
KB_DEV real kb_lg2(real x) { return log2(x); }
KB_DEV real kb_sqrt(real x) { return sqrt(x); }
KB_DEV real kb_rsqrt(real x) { return (x == 0.0) ? 0.0 : rsqrt(x); }
KB_DEV real kb_sin(real x) { return sin(x); }
KB_DEV real kb_cos(real x) { return cos(x); }
KB_DEV real kb_asin(real x) { return asin(x); }
KB_DEV real kb_acos(real x) { return acos(x); }
KB_DEV real kb_atan(real x) { return atan(x); }
KB_DEV real kb_atan2(real a, real b) { return atan2(a, b); }
// reference quirk: Powf is the C float function powf for every dtype (keopscore/utils/math_functions.py:97-99)
KB_DEV real kb_powf(real a, real b) { return (double)powf((float)a, (float)b); }
KB_DEV real kb_abs(real x) { return fabs(x); }
KB_DEV real kb_floor(real x) { return floor(x); }
KB_DEV real kb_round0(real x) { return round(x); }
KB_DEV real kb_fmax(real a, real b) { return fmax(a, b); }
KB_DEV real kb_fmin(real a, real b) { return fmin(a, b); }
#define KB_INF (__longlong_as_double(0x7ff0000000000000LL))
#else
#if KB_FAST
KB_DEV real kb_rcp(real x) { real r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KB_DEV real kb_div(real a, real b) { return __fdividef(a, b); }
KB_DEV real kb_ex2(real x) { real r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KB_DEV real kb_exp(real x) { return kb_ex2(x * 1.4426950408889634f); }
KB_DEV real kb_lg2(real x) { real r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KB_DEV real kb_log(real x) { return kb_lg2(x) * 0.6931471805599453f; }
KB_DEV real kb_sqrt(real x) { real r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
KB_DEV real kb_rsqrt(real x) { real r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return (x == 0.0f) ? 0.0f : r; }
KB_DEV real kb_sin(real x) { return __sinf(x); }
KB_DEV real kb_cos(real x) { return __cosf(x); }
KB_DEV real kb_powf(real a, real b) { return __powf(a, b); }
#else
KB_DEV real kb_rcp(real x) { return 1.0f / x; }
KB_DEV real kb_div(real a, real b) { return a / b; }
KB_DEV real kb_ex2(real x) { return exp2f(x); }
KB_DEV real kb_exp(real x) { return expf(x); }
KB_DEV real kb_lg2(real x) { return log2f(x); }
KB_DEV real kb_log(real x) { return logf(x); }
KB_DEV real kb_sqrt(real x) { return sqrtf(x); }
KB_DEV real kb_rsqrt(real x) { return (x == 0.0f) ? 0.0f : rsqrtf(x); }
KB_DEV real kb_sin(real x) { return sinf(x); }
KB_DEV real kb_cos(real x) { return cosf(x); }
KB_DEV real kb_powf(real a, real b) { return powf(a, b); }
#endif
KB_DEV real kb_asin(real x) { return asinf(x); }
KB_DEV real kb_acos(real x) { return acosf(x); }
KB_DEV real kb_atan(real x) { return atanf(x); }
KB_DEV real kb_atan2(real a, real b) { return atan2f(a, b); }
KB_DEV real kb_abs(real x) { return fabsf(x); }
KB_DEV real kb_floor(real x) { return floorf(x); }
KB_DEV real kb_round0(real x) { return roundf(x); }
KB_DEV real kb_fmax(real a, real b) { return fmaxf(a, b); }
KB_DEV real kb_fmin(real a, real b) { return fminf(a, b); }
#define KB_INF (__int_as_float(0x7f800000))
#endif

// correctly rounded square root whatever the fast-math setting (hoisted scale factors), and ex2 of the negated argument:
// the negation is free as a source modifier of MUFU.EX2, whereas the packed FMUL2 / FFMA2 have none
#if KB_F64
KB_DEV real kb_sqrt_rn(real x) { return sqrt(x); }
KB_DEV real kb_ex2n(real x) { return exp2(-x); }
#else
KB_DEV real kb_sqrt_rn(real x) { return (real)sqrt((double)x); }
#if KB_FAST
// (neg.f32 without .ftz: ptxas folds it into the MUFU source modifier; neg.ftz.f32 after a packed producer costs an FADD)
KB_DEV real kb_ex2n(real x) { real r; asm("{ .reg .f32 t; neg.f32 t, %1; ex2.approx.ftz.f32 %0, t; }" : "=f"(r) : "f"(x)); return r; }
#else
KB_DEV real kb_ex2n(real x) { return exp2f(-x); }
#endif
#endif

// smallest representable value above s (s itself when it is +inf or NaN), and an atomic minimum on reals (ordered-integer
// trick: non-negative floats order like signed ints, negative ones like reversed unsigned ints): the thresholds that the
// column splits of a K-min reduction share through global memory
#if KB_F64
KB_DEV real kb_next_up(real s) { return (s < KB_INF) ? nextafter(s, KB_INF) : s; }
KB_DEV void kb_atomic_min(real* addr, real v) {
  if (v >= 0.0) atomicMin(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
  else atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
#else
KB_DEV real kb_next_up(real s) { return (s < KB_INF) ? nextafterf(s, KB_INF) : s; }
KB_DEV void kb_atomic_min(real* addr, real v) {
  if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
#endif

// semantics of the comparison / selection family follow keopscore/utils/math_functions.py:112-150
KB_DEV real kb_sign(real x) { return (x > (real)0) ? (real)1 : ((x < (real)0) ? (real)-1 : (real)0); }
KB_DEV real kb_step(real x) { return (x < (real)0) ? (real)0 : (real)1; }
KB_DEV real kb_relu(real x) { return (x < (real)0) ? (real)0 : x; }
KB_DEV real kb_xlogx(real x) { return (x != (real)0) ? x * kb_log(x) : (real)0; }
KB_DEV real kb_sinxdivx(real x) { return (x != (real)0) ? kb_sin(x) / x : (real)1; }
KB_DEV real kb_lt(real a, real b) { return (a < b) ? (real)1 : (real)0; }
KB_DEV real kb_le(real a, real b) { return (a <= b) ? (real)1 : (real)0; }
KB_DEV real kb_eq(real a, real b) { return (a == b) ? (real)1 : (real)0; }
KB_DEV real kb_ne(real a, real b) { return (a != b) ? (real)1 : (real)0; }
KB_DEV real kb_sel_ge0(real c, real a, real b) { return (c >= (real)0) ? a : b; }
KB_DEV real kb_clamp(real x, real a, real b) { return (x < a) ? a : ((x > b) ? b : x); }
KB_DEV real kb_inrange(real x, real a, real b) { return (x < a) ? (real)0 : ((x > b) ? (real)0 : (real)1); }
KB_DEV real kb_mod(real x, real n, real d) { return x - n * kb_floor((x - d) / n); }

#define KB_LIFT1(name) \
  KB_DEV V name##_v(const V& a) { V r; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) r.l[q] = name(a.l[q]); return r; }
#define KB_LIFT2(name) \
  template <class A, class B> KB_DEV V name##_v(const A& a, const B& b) { V r; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) r.l[q] = name(kb_ln(a, q), kb_ln(b, q)); return r; }
#define KB_LIFT3(name) \
  template <class A, class B, class C> KB_DEV V name##_v(const A& a, const B& b, const C& c) { V r; KB_UNROLL_ALL for (int q = 0; q < KB_R; q++) r.l[q] = name(kb_ln(a, q), kb_ln(b, q), kb_ln(c, q)); return r; }

KB_LIFT1(kb_rcp) KB_LIFT1(kb_exp) KB_LIFT1(kb_ex2) KB_LIFT1(kb_log) KB_LIFT1(kb_lg2) KB_LIFT1(kb_sqrt)
KB_LIFT1(kb_rsqrt) KB_LIFT1(kb_sin) KB_LIFT1(kb_cos) KB_LIFT1(kb_asin) KB_LIFT1(kb_acos) KB_LIFT1(kb_atan)
KB_LIFT1(kb_abs) KB_LIFT1(kb_floor) KB_LIFT1(kb_round0) KB_LIFT1(kb_sign) KB_LIFT1(kb_step) KB_LIFT1(kb_relu)
KB_LIFT1(kb_xlogx) KB_LIFT1(kb_sinxdivx) KB_LIFT1(kb_ex2n) KB_LIFT1(kb_sqrt_rn)
KB_DEV real kb_keep(real x) { return x; }          // identity: marks a hoisted value the emitter must not re-fuse
KB_DEV V kb_keep_v(const V& a) { return a; }
KB_LIFT2(kb_div) KB_LIFT2(kb_atan2) KB_LIFT2(kb_powf) KB_LIFT2(kb_fmax) KB_LIFT2(kb_fmin) KB_LIFT2(kb_lt)
KB_LIFT2(kb_le) KB_LIFT2(kb_eq) KB_LIFT2(kb_ne)
KB_LIFT3(kb_sel_ge0) KB_LIFT3(kb_clamp) KB_LIFT3(kb_inrange) KB_LIFT3(kb_mod)

// ---------------------------------------------------------------------------------------------
// asynchronous staging of the streamed tile (LDGSTS): one element per request so that any row
// width and any 4-/8-byte aligned base pointer is accepted; zero-fill beyond the valid range.
// ---------------------------------------------------------------------------------------------
KB_DEV void kb_cp_async(real* smem_dst, const real* gmem_src, bool valid) {
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  int nbytes = valid ? (int)sizeof(real) : 0;
#if KB_F64
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gmem_src), "r"(nbytes) : "memory");
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gmem_src), "r"(nbytes) : "memory");
#endif
}
KB_DEV void kb_cp_async16(real* smem_dst, const real* gmem_src, bool valid) {
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  int nbytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(nbytes) : "memory");
}
KB_DEV void kb_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
KB_DEV void kb_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// K-min insertion: (val, idx)[0..K-1] sorted ascending; strict '<' keeps the earlier index ahead
// on ties (reference: KMin_ArgKMin_Reduction.py:71-100).  Static indexing only -> registers.
// ---------------------------------------------------------------------------------------------
template <int K, class IDX>
KB_DEV void kb_kmin_insert(real (&val)[K], IDX (&idx)[K], real f, IDX j) {
  KB_UNROLL_ALL
  for (int l = K - 1; l > 0; --l) {
    if (f < val[l - 1]) { val[l] = val[l - 1]; idx[l] = idx[l - 1]; }
    else if (f < val[l]) { val[l] = f; idx[l] = j; }
  }
  if (f < val[0]) { val[0] = f; idx[0] = j; }
}

// Same insertion with (value, index) lexicographic order: used when partial lists over *interleaved* column sets are merged
// (chunked kernel), where "earlier list wins ties" would not mean "smaller index wins ties".
template <int K>
KB_DEV void kb_kmin_insert_tie(real (&val)[K], real (&idx)[K], real f, real j) {
  KB_UNROLL_ALL
  for (int l = K - 1; l > 0; --l) {
    const bool before_prev = (f < val[l - 1]) || (f == val[l - 1] && j < idx[l - 1]);
    const bool before_here = (f < val[l]) || (f == val[l] && j < idx[l]);
    if (before_prev) { val[l] = val[l - 1]; idx[l] = idx[l - 1]; }
    else if (before_here) { val[l] = f; idx[l] = j; }
  }
  if ((f < val[0]) || (f == val[0] && j < idx[0])) { val[0] = f; idx[0] = j; }
}

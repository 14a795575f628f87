// common.cuh — arithmetic policy + kernel parameter block shared by the TPI and GPI kernels.
//
// Arithmetic contract:
//   every translation unit is compiled with -fmad=false, so plain `a*b + c` is NEVER contracted;
//   STRICT mode uses plain operators only  -> bit-identical to the pinned reference build
//   FAST   mode calls fma()/fmaf() explicitly in dot products and axpy-like updates (same order of terms).
// Division and square root are IEEE (no --use_fast_math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tmpc {

template <bool FAST, typename T>
__device__ __forceinline__ T mac(T acc, T a, T b) {  // acc + a*b
    if constexpr (FAST) {
        return fma(a, b, acc);
    } else {
        return acc + a * b;
    }
}
template <bool FAST, typename T>
__device__ __forceinline__ T nmac(T acc, T a, T b) {  // acc - a*b
    if constexpr (FAST) {
        return fma(-a, b, acc);
    } else {
        return acc - a * b;
    }
}
// ---- several ascending-k dot products against the same vector -------------------------------------------------
// out[r] = M[r][0]*v[0]; out[r] = out[r] + M[r][k]*v[k]   (k ascending; STRICT: separate multiply and add).
// fp32: rows are processed in pairs with PACKED adds (add.rn.f32x2 -> FADD2: two IEEE fp32 additions per issued
// instruction, bit-identical to two FADDs) while the multiplies stay scalar FMULs, so ptxas cannot contract them
// into FFMA2 (it fuses mul.f32x2+add.f32x2 even under -fmad=false, see profiles/r01_f32x2_microbench.txt).
// FAST fp32 uses fma.rn.f32x2 (FFMA2).  fp64 has no packed form.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// element-wise o = a + b / o = a - b on short vectors.  fp32, even length: packed FADD2 (two IEEE additions per issued
// instruction, bit-identical to two FADDs); anything else: scalar.
template <typename T, int NV>
__device__ __forceinline__ void vadd(const T (&a)[NV], const T (&b)[NV], T (&o)[NV]) {
    constexpr int NP = (sizeof(T) == 4) ? NV / 2 * 2 : 0;  // elements handled as packed pairs
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int e = 0; e < NP; e += 2) {
            float lo, hi;
            unpack2(add2(pack2((float)a[e], (float)a[e + 1]), pack2((float)b[e], (float)b[e + 1])), lo, hi);
            o[e] = lo;
            o[e + 1] = hi;
        }
    }
#pragma unroll
    for (int e = NP; e < NV; ++e) o[e] = a[e] + b[e];
}
template <typename T, int NV>
__device__ __forceinline__ void vsub(const T (&a)[NV], const T (&b)[NV], T (&o)[NV]) {
    constexpr int NP = (sizeof(T) == 4) ? NV / 2 * 2 : 0;
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int e = 0; e < NP; e += 2) {
            float lo, hi;
            unpack2(sub2(pack2((float)a[e], (float)a[e + 1]), pack2((float)b[e], (float)b[e + 1])), lo, hi);
            o[e] = lo;
            o[e + 1] = hi;
        }
    }
#pragma unroll
    for (int e = NP; e < NV; ++e) o[e] = a[e] - b[e];
}

template <bool FAST, int NR, int NE>
__device__ __forceinline__ void dots(const float (&M)[NR][NE], const float (&v)[NE], float (&out)[NR]) {
#pragma unroll
    for (int r = 0; r + 1 < NR; r += 2) {
        unsigned long long acc = pack2(M[r][0] * v[0], M[r + 1][0] * v[0]);
#pragma unroll
        for (int k = 1; k < NE; ++k) {
            if constexpr (FAST) {
                acc = fma2(pack2(M[r][k], M[r + 1][k]), pack2(v[k], v[k]), acc);
            } else {
                acc = add2(acc, pack2(M[r][k] * v[k], M[r + 1][k] * v[k]));
            }
        }
        unpack2(acc, out[r], out[r + 1]);
    }
    if constexpr (NR % 2 == 1) {
        float s = M[NR - 1][0] * v[0];
#pragma unroll
        for (int k = 1; k < NE; ++k) s = mac<FAST>(s, M[NR - 1][k], v[k]);
        out[NR - 1] = s;
    }
}
template <bool FAST, int NR, int NE>
__device__ __forceinline__ void dots(const double (&M)[NR][NE], const double (&v)[NE], double (&out)[NR]) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double s = M[r][0] * v[0];
#pragma unroll
        for (int k = 1; k < NE; ++k) s = mac<FAST>(s, M[r][k], v[k]);
        out[r] = s;
    }
}

// same, with the coefficients supplied by a callable coef(r, k) (e.g. constant-bank matrix entries in the TPI kernel)
template <bool FAST, int NR, int NE, typename F>
__device__ __forceinline__ void dots_f(F coef, const float (&v)[NE], float (&out)[NR]) {
#pragma unroll
    for (int r = 0; r + 1 < NR; r += 2) {
        unsigned long long acc = pack2(coef(r, 0) * v[0], coef(r + 1, 0) * v[0]);
#pragma unroll
        for (int k = 1; k < NE; ++k) {
            if constexpr (FAST) {
                acc = fma2(pack2(coef(r, k), coef(r + 1, k)), pack2(v[k], v[k]), acc);
            } else {
                acc = add2(acc, pack2(coef(r, k) * v[k], coef(r + 1, k) * v[k]));
            }
        }
        unpack2(acc, out[r], out[r + 1]);
    }
    if constexpr (NR % 2 == 1) {
        float s = coef(NR - 1, 0) * v[0];
#pragma unroll
        for (int k = 1; k < NE; ++k) s = mac<FAST>(s, coef(NR - 1, k), v[k]);
        out[NR - 1] = s;
    }
}
template <bool FAST, int NR, int NE, typename F>
__device__ __forceinline__ void dots_f(F coef, const double (&v)[NE], double (&out)[NR]) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double s = coef(r, 0) * v[0];
#pragma unroll
        for (int k = 1; k < NE; ++k) s = mac<FAST>(s, coef(r, k), v[k]);
        out[r] = s;
    }
}

template <typename T>
__device__ __forceinline__ T tabs(T a) {
    return a < T(0) ? -a : a;
}
__device__ __forceinline__ float tabs(float a) { return fabsf(a); }
__device__ __forceinline__ double tabs(double a) { return fabs(a); }
__device__ __forceinline__ float tsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double tsqrt(double a) { return sqrt(a); }

// Eigen's x_max.cwiseMin(x_min.cwiseMax(v)) as executed: m = (lo<v)?v:lo ; r = (m<hi)?m:hi
template <typename T>
__device__ __forceinline__ T clamp_ref(T v, T lo, T hi) {
    T m = (lo < v) ? v : lo;
    return (m < hi) ? m : hi;
}

// project_soc for a 3-vector exactly as the reference executes it (admm.cpp:39-60):
// mu and the norm are narrowed to float even when T = double.
template <typename T>
__device__ __forceinline__ void project_soc3(T &s0, T &s1, T &s2, T mu_T) {
    const float mu = (float)mu_T;
    const T u0 = s2 * (T)mu;
    T sq = s0 * s0;
    sq = sq + s1 * s1;
    const float a = (float)tsqrt(sq);
    if ((T)a <= -u0) {
        s0 = T(0);
        s1 = T(0);
        s2 = T(0);
    } else if ((T)a <= u0) {
        // inside the cone: unchanged
    } else {
        const T third = (T)(a / mu);  // float division (admm.cpp:54)
        const T c = T(0.5) * (T(1) + u0 / (T)a);
        s0 = c * s0;
        s1 = c * s1;
        s2 = c * third;
    }
}

// branch-free twin of project_soc3 (same operations, same results): every lane of a warp projects a different cone, so
// the three cases are computed and selected instead of branched on
template <typename T>
__device__ __forceinline__ void project_soc3_sel(T &s0, T &s1, T &s2, T mu_T) {
    const float mu = (float)mu_T;
    const T u0 = s2 * (T)mu;
    T sq = s0 * s0;
    sq = sq + s1 * s1;
    const float a = (float)tsqrt(sq);
    const T ad = (T)a;
    const bool below = ad <= -u0, inside = ad <= u0;
    const T third = (T)(a / mu);
    const T c = T(0.5) * (T(1) + u0 / ad);
    const T r0 = c * s0, r1 = c * s1, r2 = c * third;
    s0 = below ? T(0) : (inside ? s0 : r0);
    s1 = below ? T(0) : (inside ? s1 : r1);
    s2 = below ? T(0) : (inside ? s2 : r2);
}

constexpr int MAX_CONES = 4;  // cones per knot point and per side held in the parameter block

// Run-time part of the streamed lane-group kernel's workspace description (gps_kernel.cuh); the record layout itself is a
// compile-time function of the shape and of the compiled-in constraint families (GpsRec).
struct GpsLayout {
    int has_b;  // the optional-output region (previous box slacks, family slacks) is allocated and maintained
};

// Kernel parameter block.  Passed by value (__grid_constant__): it lives in the constant bank, so with
// compile-time indices the matrix entries become immediate constant operands of the FMA instructions.
// All matrices are column-major copies of the reference's cache / workspace (types.hpp:43-51,186-190).
template <typename T, int NX, int NU>
struct KParams {
    T A[NX * NX];
    T Bm[NX * NU];
    T f[NX];
    T Qd[NX];
    T Rd[NU];
    T Kinf[NU * NX];
    T Pinf[NX * NX];
    T Quu[NU * NU];
    T AmBKt[NX * NX];
    T APf[NX];
    T BPf[NU];
    T rho, pri_tol, dua_tol;
    int N, max_iter, check_termination;
    int en_state_bound, en_input_bound;
    int soc_x, soc_u;    // en_*_soc && num cones > 0  (admm.cpp:102,107)
    int ncx, ncu;        // cone loops run under en_*_soc alone (admm.cpp:112,125)
    int lin_x, lin_u, nlx, nlu;
    int tvl_x, tvl_u, ntvx, ntvu;
    int cone_x_start[MAX_CONES], cone_u_start[MAX_CONES];
    T cone_x_mu[MAX_CONES], cone_u_mu[MAX_CONES];
    int64_t B;     // instances in this launch
    int64_t Bpad;  // workspace stride (instances, multiple of 32)
    int cold;
    int bounds_tv;      // bounds vary along the horizon (else row 0 is used for every k)
    T xlo[NX], xhi[NX], ulo[NU], uhi[NU];  // time-invariant bounds (column 0); (-inf, +inf) when the bound is disabled
    const T *Pinf_g;    // Pinf in global memory (column-major), GPI terminal cost
    int xref_pi, uref_pi;
    // inputs (user layout, instance-major)
    const T *x0, *Xref, *Uref;
    // bounds (device, column-major nx x N etc.)
    const T *x_min, *x_max, *u_min, *u_max;
    // hyperplanes (device)
    const T *Alin_x, *blin_x, *Alin_u, *blin_u;
    const T *tv_Alin_x, *tv_blin_x, *tv_Alin_u, *tv_blin_u;
    // warm-start state in/out (user layout), any may be null
    T *s_x, *s_u, *s_v, *s_z, *s_vnew, *s_znew, *s_g, *s_y;
    T *s_vcnew, *s_zcnew, *s_gc, *s_yc, *s_vlnew, *s_zlnew, *s_gl, *s_yl;
    T *s_vlnew_tv, *s_zlnew_tv, *s_gl_tv, *s_yl_tv;
    // outputs
    T *sol_x, *sol_u;  // may be null
    int32_t *iter, *solved;
    T *residuals;
    T *u0;  // [B][nu] first rollout input, may be null
    const T *models;  // GPI: per-instance cache blobs [B][blob+1] (A,B,f,Qd,Rd,Kinf,Pinf,Quu,AmBKt,APf,BPf,rho) or null
    T *gpi_vscratch;  // GPI: [B][N][L][PVP] copy of the previous primal pack (work->v / work->z) while v,z are persisted
    T *gps_ws;        // GPS: [resident warps][N][gps.rec] streamed state records
    GpsLayout gps;
    // TPI workspace (structure-of-arrays, 16-byte vectors, [k][vec][Bpad])
    void *w_v[2], *w_z[2], *w_g, *w_y, *w_d;
    void *w_vc, *w_zc, *w_gc, *w_yc, *w_vl, *w_zl, *w_gl, *w_yl, *w_vlt, *w_zlt, *w_glt, *w_ylt;
};

}  // namespace tmpc

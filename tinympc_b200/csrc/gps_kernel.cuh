// gps_kernel.cuh — lane-group-per-instance batched ADMM solve with the N-indexed state STREAMED through a per-slot
// workspace in global memory (L2 / HBM) behind a cp.async ring in shared memory ("GPS").
//
// Same lane mapping as the on-chip kernel (gpi_kernel.cuh): L lanes own one MPC instance, lane l owns the state rows
// [l*RX, (l+1)*RX) and the input rows [l*RU, (l+1)*RU); its rows of AmBKt / B^T / A / Kinf / Kinf^T / B / Quu_inv live
// in registers (staged once per CTA by a TMA bulk copy), every mat-vec is RX+RU ascending-k dot products per lane and
// the fresh vector is all-gathered inside the lane group through shared memory.  What differs:
//   * every constraint family of the reference is covered: box (admm.cpp:85-98), second-order cones (project_soc,
//     admm.cpp:39-60,102-135), static and time-varying hyperplanes (admm.cpp:70-73,138-211) with their cost / dual
//     twins (admm.cpp:228-255,268-303), in fp32 and fp64;
//   * the per-instance state does not have to fit on chip (rocket landing, N = 100, fp64: 31 KB per instance): it lives
//     in a workspace indexed by RESIDENT SLOT (not by instance), one record per (warp, knot point) laid out
//     [field][instance of the warp][row], so that every field access of a warp is one contiguous run of bytes.  Each
//     lane moves only its own rows: cp.async (4/8/16-byte chunks) into a private slice of a shared-memory ring
//     `dist` steps ahead of their use, plain vector stores on the way out.  A lane only ever re-reads what it wrote
//     itself, so program order is all the ordering the ring needs (no barriers, no drain at the sweep turnarounds);
//   * the linear cost of the NEXT iteration (q_k, r_k, p_{N-1}; update_linear_cost, admm.cpp:262-304) is evaluated
//     in the forward sweep, where the fresh slack / dual values are in registers, and stored: the backward sweep
//     reads q, r (nx+nu values per knot point) instead of every slack / dual pair (up to 8 (nx+nu)).
// The kernel is persistent (one CTA per SM); slots are refilled from a global atomic queue as instances terminate
// (per-instance termination, admm.cpp:310-328).  Reference semantics: tiny_solve -> solve (admm.cpp:331-455).
#pragma once
#include "tpi_kernel.cuh"
#include "gpi_kernel.cuh"

namespace tmpc {

constexpr int gps_gcd(int a, int b) { return b == 0 ? a : gps_gcd(b, a % b); }

template <int NX, int NU, int L, int ES, bool EXT>
struct GpsCfg {
    static constexpr int RX = (NX + L - 1) / L;
    static constexpr int RU = (NU + L - 1) / L;
    static constexpr int IPW = 32 / L;
    static constexpr int W = 16 / ES;
    static constexpr int NXP = (L * RX + W - 1) / W * W;
    static constexpr int NUP = (L * RU + W - 1) / W * W;
    static constexpr int GBUF = IPW * (NXP > NUP ? NXP : NUP);
    // chunk sizes (bytes) of a lane's piece: global side (limited by the row pitch of an instance) and shared side
    static constexpr int CX = gps_gcd(16, gps_gcd(NX * ES, RX * ES));
    static constexpr int CU = gps_gcd(16, gps_gcd(NU * ES, RU * ES));
    static constexpr int SX = gps_gcd(16, RX * ES);
    static constexpr int SU = gps_gcd(16, RU * ES);
    // ring piece slots.  state-shaped: 0 vnew (backward: q), 1 g, 2 xref, 3.. family duals
    //                    input-shaped: 0 d (backward: r), 1 znew, 2 y, 3 uref, 4.. family duals
    static constexpr int XP = EXT ? 6 : 3;
    static constexpr int UP = EXT ? 7 : 4;
    static constexpr int PXB = 32 * RX * ES;  // bytes of one state-shaped piece slot (32 lanes)
    static constexpr int PUB = 32 * RU * ES;
    static constexpr int STAGE_BYTES = XP * PXB + UP * PUB;
    static constexpr int MAT_REGS = RX * (2 * NX + 2 * NU + 3) + RU * (2 * NX + NU + 2);
    static constexpr bool ok = (NX % RX == 0) && (NU % RU == 0) && (MAT_REGS * (ES / 4) <= 150);
    __host__ __device__ static constexpr size_t warp_bytes(int stages) { return (size_t)GBUF * ES + (size_t)stages * STAGE_BYTES; }
};

// smallest lane-group width whose matrix rows fit in registers (0 = none)
template <typename T, int NX, int NU>
constexpr int gps_pick_L() {
    if (GpsCfg<NX, NU, 4, (int)sizeof(T), true>::ok) return 4;
    if (GpsCfg<NX, NU, 8, (int)sizeof(T), true>::ok) return 8;
    if (GpsCfg<NX, NU, 16, (int)sizeof(T), true>::ok) return 16;
    return 0;
}

constexpr int GPS_MAX_WARPS = 8;

// ---- cp.async (per-thread asynchronous global -> shared copies) ----
template <int BYTES>
__device__ __forceinline__ void cp_async(unsigned dst, const void *src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async moves 4, 8 or 16 bytes");
    if constexpr (BYTES == 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else if constexpr (BYTES == 8) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
    }
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait(int pending) {  // warp-uniform argument
    if (pending <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else if (pending == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else asm volatile("cp.async.wait_group 3;" ::: "memory");
}
// a lane's piece of PB bytes in chunks of CB bytes
template <int PB, int CB>
__device__ __forceinline__ void cp_piece(unsigned dst, const void *src) {
#pragma unroll
    for (int c = 0; c < PB / CB; ++c) cp_async<CB>(dst + (unsigned)(c * CB), reinterpret_cast<const char *>(src) + c * CB);
}

// ---- chunked piece moves between registers and shared / global memory ----
__device__ __forceinline__ void lds_chunk(unsigned a, float (&v)[1]) { v[0] = lds(a, 0.f); }
__device__ __forceinline__ void lds_chunk(unsigned a, float (&v)[2]) {
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a));
}
__device__ __forceinline__ void lds_chunk(unsigned a, float (&v)[4]) { ldsv(a, v); }
__device__ __forceinline__ void lds_chunk(unsigned a, double (&v)[1]) { v[0] = lds(a, 0.0); }
__device__ __forceinline__ void lds_chunk(unsigned a, double (&v)[2]) { ldsv(a, v); }
__device__ __forceinline__ void stg_chunk(float *p, const float (&v)[1]) { *p = v[0]; }
__device__ __forceinline__ void stg_chunk(float *p, const float (&v)[2]) { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); }
__device__ __forceinline__ void stg_chunk(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void stg_chunk(double *p, const double (&v)[1]) { *p = v[0]; }
__device__ __forceinline__ void stg_chunk(double *p, const double (&v)[2]) { *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]); }

template <typename T, int R, int CB>
__device__ __forceinline__ void lds_piece(unsigned a, T (&v)[R]) {
    constexpr int E = CB / (int)sizeof(T);
#pragma unroll
    for (int c = 0; c < R / E; ++c) {
        T t[E];
        lds_chunk(a + (unsigned)(c * CB), t);
#pragma unroll
        for (int e = 0; e < E; ++e) v[c * E + e] = t[e];
    }
}
template <typename T, int R, int CB>
__device__ __forceinline__ void stg_piece(T *p, const T (&v)[R]) {
    constexpr int E = CB / (int)sizeof(T);
#pragma unroll
    for (int c = 0; c < R / E; ++c) {
        T t[E];
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = v[c * E + e];
        stg_chunk(p + c * E, t);
    }
}
template <typename T, int R, int CB>
__device__ __forceinline__ void ldg_piece(const T *p, T (&v)[R]) {  // plain global loads of a piece (write-back paths)
#pragma unroll
    for (int e = 0; e < R; ++e) v[e] = p[e];
}

// own rows of a vector every lane of the group holds in full: out[a] = full[l*R + a] without dynamic register indexing
template <typename T, int NE, int R, int L>
__device__ __forceinline__ void extract_own(const T (&full)[NE], int l, T (&out)[R]) {
#pragma unroll
    for (int a = 0; a < R; ++a) {
        T v = T(0);
#pragma unroll
        for (int g = 0; g < L; ++g)
            if (g * R + a < NE) v = (l == g) ? full[g * R + a] : v;
        out[a] = v;
    }
}

template <typename T, int NX, int NU, int L, bool FAST, bool EXT>
__global__ void __launch_bounds__(GPS_MAX_WARPS * 32, 1)
    gps_solve_kernel(const __grid_constant__ KParams<T, NX, NU> P, const T *__restrict__ gmat, unsigned long long *queue) {
    using Cfg = GpsCfg<NX, NU, L, (int)sizeof(T), EXT>;
    constexpr int RX = Cfg::RX, RU = Cfg::RU, W = Cfg::W, NXP = Cfg::NXP, NUP = Cfg::NUP;
    constexpr int CX = Cfg::CX, CU = Cfg::CU, SX = Cfg::SX, SU = Cfg::SU;
    constexpr unsigned ES = (unsigned)sizeof(T);
    constexpr unsigned PXB = Cfg::PXB, PUB = Cfg::PUB, STAGE = Cfg::STAGE_BYTES, XPB = Cfg::XP * Cfg::PXB;
    static_assert(Cfg::ok, "lane mapping not available for this shape");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = P.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int l = lane % L, slot = lane / L;
    const GpsLayout &LY = P.gps;
    const int D = LY.dist, S = D + 1;
    const int rec = LY.rec;
    const T rho = P.rho;

    // ---- stage the cache blob into shared memory with one TMA bulk copy per CTA, pull this lane's rows into registers
    constexpr int OFF_A = 0, OFF_B = OFF_A + NX * NX, OFF_F = OFF_B + NX * NU, OFF_QD = OFF_F + NX, OFF_RD = OFF_QD + NX,
                  OFF_K = OFF_RD + NU, OFF_PINF = OFF_K + NU * NX, OFF_QUU = OFF_PINF + NX * NX,
                  OFF_AMBKT = OFF_QUU + NU * NU, OFF_APF = OFF_AMBKT + NX * NX, OFF_BPF = OFF_APF + NX,
                  BLOB = OFF_BPF + NU;
    constexpr unsigned BLOB_BYTES = (unsigned)(((BLOB * sizeof(T) + 15) / 16) * 16);
    T *stage = reinterpret_cast<T *>(smem_raw);
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BLOB_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(stage)),
                     "l"(gmat), "r"(BLOB_BYTES), "r"(mb)
                     : "memory");
    }
    __syncthreads();
    {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(mb)
                : "memory");
        }
    }
    const bool xvl = l * RX < NX, uvl = l * RU < NU;  // does this lane own real rows (else padding rows: zeros)
    T mS1b[RX + RU][NX], mS1f[RX + RU][NX];
    T mKt[RX][NU], mB[RX][NU], vQd[RX], vAPf[RX], vf[RX];
    T mQuu[RU][NU], vRd[RU], vBPf[RU];
    {
        const T *src = stage;
#pragma unroll
        for (int a = 0; a < RX; ++a) {
            const int ii = xvl ? l * RX + a : 0;
#pragma unroll
            for (int m = 0; m < NX; ++m) {
                mS1b[a][m] = xvl ? src[OFF_AMBKT + ii + NX * m] : T(0);
                mS1f[a][m] = xvl ? src[OFF_A + ii + NX * m] : T(0);
            }
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                mKt[a][j] = xvl ? src[OFF_K + j + NU * ii] : T(0);
                mB[a][j] = xvl ? src[OFF_B + ii + NX * j] : T(0);
            }
            vQd[a] = xvl ? src[OFF_QD + ii] : T(0);
            vAPf[a] = xvl ? src[OFF_APF + ii] : T(0);
            vf[a] = xvl ? src[OFF_F + ii] : T(0);
        }
#pragma unroll
        for (int b = 0; b < RU; ++b) {
            const int jj = uvl ? l * RU + b : 0;
#pragma unroll
            for (int m = 0; m < NX; ++m) {
                mS1b[RX + b][m] = uvl ? src[OFF_B + m + NX * jj] : T(0);
                mS1f[RX + b][m] = uvl ? src[OFF_K + jj + NU * m] : T(0);
            }
#pragma unroll
            for (int m = 0; m < NU; ++m) mQuu[b][m] = uvl ? src[OFF_QUU + jj + NU * m] : T(0);
            vRd[b] = uvl ? src[OFF_RD + jj] : T(0);
            vBPf[b] = uvl ? src[OFF_BPF + jj] : T(0);
        }
    }
    __syncthreads();  // the staging area is reused below

    // ---- shared memory of this warp: gather scratch + cp.async ring (S stages) ----
    const unsigned wbytes = (unsigned)Cfg::warp_bytes(S);
    const unsigned aGB = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)warp * wbytes;
    const unsigned aRing = aGB + (unsigned)Cfg::GBUF * ES;
    const unsigned aXl = aRing + (unsigned)lane * RX * ES;        // + stage*STAGE + piece*PXB
    const unsigned aUl = aRing + XPB + (unsigned)lane * RU * ES;  // + stage*STAGE + piece*PUB
    // padding lanes never receive data: their ring slices stay zero
    for (unsigned o = (unsigned)lane * 16u; o < (unsigned)S * STAGE; o += 32u * 16u)
        asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(aRing + o), "f"(0.f) : "memory");
    __syncwarp();

    auto gather_x = [&](const T (&own)[RX], T (&full)[NX]) {
        __syncwarp();
#pragma unroll
        for (int a = 0; a < RX; ++a) sts(aGB + (unsigned)(slot * NXP + l * RX + a) * ES, own[a]);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < NXP / W; ++c) {
            T t[W];
            ldsv(aGB + (unsigned)(slot * NXP + c * W) * ES, t);
#pragma unroll
            for (int e = 0; e < W; ++e)
                if (c * W + e < NX) full[c * W + e] = t[e];
        }
    };
    auto gather_u = [&](const T (&own)[RU], T (&full)[NU]) {
        __syncwarp();
#pragma unroll
        for (int b = 0; b < RU; ++b) sts(aGB + (unsigned)(slot * NUP + l * RU + b) * ES, own[b]);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < NUP / W; ++c) {
            T t[W];
            ldsv(aGB + (unsigned)(slot * NUP + c * W) * ES, t);
#pragma unroll
            for (int e = 0; e < W; ++e)
                if (c * W + e < NU) full[c * W + e] = t[e];
        }
    };

    // ---- streamed state of this warp: records [k][field][instance of the warp][row] ----
    T *const wsw = P.gps_ws + (int64_t)(blockIdx.x * nwarps + warp) * N * rec;
    T *const gx = wsw + slot * NX + l * RX;  // this lane's rows inside a state-shaped field of record 0
    T *const gu = wsw + slot * NU + l * RU;

    const bool cold = P.cold != 0;
    const bool tvb = P.bounds_tv != 0;
    const bool enx = P.en_state_bound != 0, enu = P.en_input_bound != 0;
    const T kInf = (T)INFINITY;
    T loX[RX], hiX[RX], loU[RU], hiU[RU];
#pragma unroll
    for (int a = 0; a < RX; ++a) {
        loX[a] = (enx && xvl) ? __ldg(P.x_min + l * RX + a) : -kInf;
        hiX[a] = (enx && xvl) ? __ldg(P.x_max + l * RX + a) : kInf;
    }
#pragma unroll
    for (int b = 0; b < RU; ++b) {
        loU[b] = (enu && uvl) ? __ldg(P.u_min + l * RU + b) : -kInf;
        hiU[b] = (enu && uvl) ? __ldg(P.u_max + l * RU + b) : kInf;
    }
    const bool keep_v = LY.vprev >= 0;
    const bool fx[3] = {EXT && P.soc_x != 0, EXT && P.lin_x != 0, EXT && P.tvl_x != 0};
    const bool fu[3] = {EXT && P.soc_u != 0, EXT && P.lin_u != 0, EXT && P.tvl_u != 0};
    const bool has_uref = P.Uref != nullptr;

    // ---- per-slot bookkeeping (identical in the L lanes of a slot) ----
    int64_t inst = -1;
    bool busy = false, want = true;
    int it = 0, solved = 0;
    T res_px = T(0), res_dx = T(0), res_pu = T(0), res_du = T(0);
    T x0o[RX], pterm[RX];
#pragma unroll
    for (int a = 0; a < RX; ++a) x0o[a] = pterm[a] = T(0);
    const T *xrefp = P.Xref + l * RX;
    const T *urefp = has_uref ? P.Uref + l * RU : P.Xref;

    // ---- projections of one gathered column (EXT) ----
    // cones (admm.cpp:102-135): the lane group's vector is in the gather scratch; every lane projects each cone's
    // three rows (project_soc, admm.cpp:39-60) and keeps the rows it owns.  Cones are pairwise disjoint (checked
    // by the host; overlapping cones run on the thread-per-instance kernel).
    auto cones_x = [&](T (&own)[RX]) {
        __syncwarp();
#pragma unroll
        for (int a = 0; a < RX; ++a) sts(aGB + (unsigned)(slot * NXP + l * RX + a) * ES, own[a]);
        __syncwarp();
        for (int c = 0; c < P.ncx; ++c) {
            const int st0 = P.cone_x_start[c];
            T s0 = lds(aGB + (unsigned)(slot * NXP + st0) * ES, T()), s1 = lds(aGB + (unsigned)(slot * NXP + st0 + 1) * ES, T()),
              s2 = lds(aGB + (unsigned)(slot * NXP + st0 + 2) * ES, T());
            project_soc3(s0, s1, s2, P.cone_x_mu[c]);
#pragma unroll
            for (int a = 0; a < RX; ++a) {
                const int i = l * RX + a;
                own[a] = (i == st0) ? s0 : (i == st0 + 1) ? s1 : (i == st0 + 2) ? s2 : own[a];
            }
        }
    };
    auto cones_u = [&](T (&own)[RU]) {
        __syncwarp();
#pragma unroll
        for (int b = 0; b < RU; ++b) sts(aGB + (unsigned)(slot * NUP + l * RU + b) * ES, own[b]);
        __syncwarp();
        for (int c = 0; c < P.ncu; ++c) {
            const int st0 = P.cone_u_start[c];
            T s0 = lds(aGB + (unsigned)(slot * NUP + st0) * ES, T()), s1 = lds(aGB + (unsigned)(slot * NUP + st0 + 1) * ES, T()),
              s2 = lds(aGB + (unsigned)(slot * NUP + st0 + 2) * ES, T());
            project_soc3(s0, s1, s2, P.cone_u_mu[c]);
#pragma unroll
            for (int b = 0; b < RU; ++b) {
                const int j = l * RU + b;
                own[b] = (j == st0) ? s0 : (j == st0 + 1) ? s1 : (j == st0 + 2) ? s2 : own[b];
            }
        }
    };
    // hyperplanes (admm.cpp:138-211): the rows are applied one after the other to the whole column, so every lane of
    // the group carries the full vector through the sequence (same arithmetic on every lane) and keeps its own rows
    auto planes_x = [&](T (&own)[RX], const T *A, int ld, int row0, int n, const T *bvec) {
        T full[NX];
        gather_x(own, full);
        project_rows<FAST, T, NX>(full, A, ld, row0, n, bvec);
        T o[RX];
        extract_own<T, NX, RX, L>(full, l, o);
#pragma unroll
        for (int a = 0; a < RX; ++a) own[a] = xvl ? o[a] : own[a];
    };
    auto planes_u = [&](T (&own)[RU], const T *A, int ld, int row0, int n, const T *bvec) {
        T full[NU];
        gather_u(own, full);
        project_rows<FAST, T, NU>(full, A, ld, row0, n, bvec);
        T o[RU];
        extract_own<T, NU, RU, L>(full, l, o);
#pragma unroll
        for (int b = 0; b < RU; ++b) own[b] = uvl ? o[b] : own[b];
    };

    // ---- ring producers: all copies of one sweep step form one cp.async group ----
    auto issue_fwd = [&](int k, int st) {
        if (k < N) {
            const unsigned bx = aXl + (unsigned)st * STAGE, bu = aUl + (unsigned)st * STAGE;
            const T *rx = gx + (int64_t)k * rec;
            const T *ru = gu + (int64_t)k * rec;
            if (xvl) {
                cp_piece<RX * ES, CX>(bx, rx + LY.vnew);
                cp_piece<RX * ES, CX>(bx + PXB, rx + LY.g);
                cp_piece<RX * ES, ES>(bx + 2 * PXB, xrefp + (int64_t)k * NX);
                if constexpr (EXT) {
#pragma unroll
                    for (int f = 0; f < 3; ++f)
                        if (fx[f]) cp_piece<RX * ES, CX>(bx + (3 + f) * PXB, rx + LY.gf[f]);
                }
            }
            if (uvl && k < N - 1) {
                cp_piece<RU * ES, CU>(bu, ru + LY.d);
                cp_piece<RU * ES, CU>(bu + PUB, ru + LY.znew);
                cp_piece<RU * ES, CU>(bu + 2 * PUB, ru + LY.y);
                if (has_uref) cp_piece<RU * ES, ES>(bu + 3 * PUB, urefp + (int64_t)k * NU);
                if constexpr (EXT) {
#pragma unroll
                    for (int f = 0; f < 3; ++f)
                        if (fu[f]) cp_piece<RU * ES, CU>(bu + (4 + f) * PUB, ru + LY.yf[f]);
                }
            }
        }
        cp_commit();
    };
    auto issue_bwd = [&](int k, int st) {
        if (k >= 0) {
            const unsigned bx = aXl + (unsigned)st * STAGE, bu = aUl + (unsigned)st * STAGE;
            if (xvl) cp_piece<RX * ES, CX>(bx, gx + (int64_t)k * rec + LY.q);
            if (uvl && k < N - 1) cp_piece<RU * ES, CU>(bu, gu + (int64_t)k * rec + LY.r);
        }
        cp_commit();
    };

    // ---- forward sweep: rollout (admm.cpp:25-32) fused with update_slack (:81-213), update_dual (:219-256), the
    // residual maxima of termination_condition (:310-328) and the NEXT iteration's update_linear_cost (:262-304) ----
    auto forward = [&](T &rpx, T &rdx, T &rpu, T &rdu) {
        T xo[RX], Xf[NX];
#pragma unroll
        for (int a = 0; a < RX; ++a) xo[a] = x0o[a];
        for (int j = 0; j < D; ++j) issue_fwd(j, j);
        gather_x(xo, Xf);
        int st = 0, sti = D % S;
        auto column = [&](int k, const bool HASU) {  // always inlined with a literal HASU
            issue_fwd(k + D, sti);
            cp_wait(D);
            const unsigned bx = aXl + (unsigned)st * STAGE, bu = aUl + (unsigned)st * STAGE;
            T vo[RX], g[RX], xr[RX];
            lds_piece<T, RX, SX>(bx, vo);
            lds_piece<T, RX, SX>(bx + PXB, g);
            lds_piece<T, RX, SX>(bx + 2 * PXB, xr);
            T t1[RX + RU], u[RU], Uf[NU];
#pragma unroll
            for (int b = 0; b < RU; ++b) u[b] = T(0);
            if (HASU) {
                T dk[RU];
                lds_piece<T, RU, SU>(bu, dk);
                dots<FAST>(mS1f, Xf, t1);  // [A x_k ; Kinf x_k]
#pragma unroll
                for (int b = 0; b < RU; ++b) u[b] = (-t1[RX + b]) - dk[b];  // u_k = -(Kinf x_k) - d_k
                gather_u(u, Uf);
            }
            if (tvb) {
#pragma unroll
                for (int a = 0; a < RX; ++a) {
                    loX[a] = (enx && xvl) ? __ldg(P.x_min + (int64_t)k * NX + l * RX + a) : loX[a];
                    hiX[a] = (enx && xvl) ? __ldg(P.x_max + (int64_t)k * NX + l * RX + a) : hiX[a];
                }
                if (HASU) {
#pragma unroll
                    for (int b = 0; b < RU; ++b) {
                        loU[b] = (enu && uvl) ? __ldg(P.u_min + (int64_t)k * NU + l * RU + b) : loU[b];
                        hiU[b] = (enu && uvl) ? __ldg(P.u_max + (int64_t)k * NU + l * RU + b) : hiU[b];
                    }
                }
            }
            T *const wx = gx + (int64_t)k * rec;
            T *const wu = gu + (int64_t)k * rec;
            {   // state column k
                T vn[RX], gn[RX], q[RX];
#pragma unroll
                for (int a = 0; a < RX; ++a) {
                    const T v = clamp_box<FAST>(xo[a] + g[a], loX[a], hiX[a]);  // vnew = clamp(x + g)
                    vn[a] = v;
                    gn[a] = (g[a] + xo[a]) - v;                                   // g += x - vnew
                    rpx = absmax(rpx, xo[a] - v);
                    rdx = absmax(rdx, vo[a] - v);
                    // k < N-1: q_k = -(xref*Q) - rho (vnew - g);  k = N-1: p_{N-1} = -(Pinf^T xref) - rho (vnew - g)
                    const T base = HASU ? -(xr[a] * vQd[a]) : pterm[a];
                    q[a] = nmac<FAST>(base, rho, v - gn[a]);
                }
                if (xvl) {
                    stg_piece<T, RX, CX>(wx + LY.vnew, vn);
                    stg_piece<T, RX, CX>(wx + LY.g, gn);
                    if (keep_v) stg_piece<T, RX, CX>(wx + LY.vprev, vo);
                }
                if constexpr (EXT) {
#pragma unroll
                    for (int f = 0; f < 3; ++f) {
                        if (fx[f]) {  // warp-uniform
                            T gf[RX], sf[RX], gfn[RX];
                            lds_piece<T, RX, SX>(bx + (3 + f) * PXB, gf);
#pragma unroll
                            for (int a = 0; a < RX; ++a) sf[a] = xo[a] + gf[a];
                            if (f == 0) cones_x(sf);
                            else if (f == 1) planes_x(sf, P.Alin_x, P.nlx, 0, P.nlx, P.blin_x);
                            else planes_x(sf, P.tv_Alin_x, P.ntvx * N, P.ntvx * k, P.ntvx, P.tv_blin_x + (int64_t)k * P.ntvx);
#pragma unroll
                            for (int a = 0; a < RX; ++a) {
                                gfn[a] = (gf[a] + xo[a]) - sf[a];
                                q[a] = nmac<FAST>(q[a], rho, sf[a] - gfn[a]);
                            }
                            if (xvl) {
                                stg_piece<T, RX, CX>(wx + LY.gf[f], gfn);
                                if (LY.vf[f] >= 0) stg_piece<T, RX, CX>(wx + LY.vf[f], sf);
                            }
                        }
                    }
                }
                if (xvl) stg_piece<T, RX, CX>(wx + LY.q, q);
            }
            if (HASU) {  // input column k and the rollout step
                T zo[RU], y[RU], ur[RU], zn[RU], yn[RU], r[RU];
                lds_piece<T, RU, SU>(bu + PUB, zo);
                lds_piece<T, RU, SU>(bu + 2 * PUB, y);
                lds_piece<T, RU, SU>(bu + 3 * PUB, ur);
#pragma unroll
                for (int b = 0; b < RU; ++b) {
                    const T z = clamp_box<FAST>(u[b] + y[b], loU[b], hiU[b]);
                    zn[b] = z;
                    yn[b] = (y[b] + u[b]) - z;
                    rpu = absmax(rpu, u[b] - z);
                    rdu = absmax(rdu, zo[b] - z);
                    const T urb = has_uref ? ur[b] : T(0);
                    r[b] = nmac<FAST>(-(urb * vRd[b]), rho, z - yn[b]);
                }
                if (uvl) {
                    stg_piece<T, RU, CU>(wu + LY.znew, zn);
                    stg_piece<T, RU, CU>(wu + LY.y, yn);
                    if (keep_v) stg_piece<T, RU, CU>(wu + LY.zprev, zo);
                }
                if constexpr (EXT) {
#pragma unroll
                    for (int f = 0; f < 3; ++f) {
                        if (fu[f]) {
                            T yf[RU], sf[RU], yfn[RU];
                            lds_piece<T, RU, SU>(bu + (4 + f) * PUB, yf);
#pragma unroll
                            for (int b = 0; b < RU; ++b) sf[b] = u[b] + yf[b];
                            if (f == 0) cones_u(sf);
                            else if (f == 1) planes_u(sf, P.Alin_u, P.nlu, 0, P.nlu, P.blin_u);
                            else planes_u(sf, P.tv_Alin_u, P.ntvu * (N - 1), P.ntvu * k, P.ntvu, P.tv_blin_u + (int64_t)k * P.ntvu);
#pragma unroll
                            for (int b = 0; b < RU; ++b) {
                                yfn[b] = (yf[b] + u[b]) - sf[b];
                                r[b] = nmac<FAST>(r[b], rho, sf[b] - yfn[b]);
                            }
                            if (uvl) {
                                stg_piece<T, RU, CU>(wu + LY.yf[f], yfn);
                                if (LY.zf[f] >= 0) stg_piece<T, RU, CU>(wu + LY.zf[f], sf);
                            }
                        }
                    }
                }
                if (uvl) stg_piece<T, RU, CU>(wu + LY.r, r);
                // x_{k+1} = (A x_k + B u_k) + f                                            (admm.cpp:30)
                T bu_[RX];
                dots<FAST>(mB, Uf, bu_);
#pragma unroll
                for (int a = 0; a < RX; ++a) xo[a] = (t1[a] + bu_[a]) + vf[a];
                gather_x(xo, Xf);
            }
            st = (st + 1 == S) ? 0 : st + 1;
            sti = (sti + 1 == S) ? 0 : sti + 1;
        };
        for (int k = 0; k < N - 1; ++k) column(k, true);
        column(N - 1, false);
    };

    // ---- backward sweep (admm.cpp:13-20) on the stored linear cost ----
    auto backward = [&]() {
        for (int j = 0; j < D; ++j) issue_bwd(N - 1 - j, j);
        int st = 0, sti = D % S;
        T po[RX], Pf[NX];
        {   // terminal cost p_{N-1}
            issue_bwd(N - 1 - D, sti);
            cp_wait(D);
            lds_piece<T, RX, SX>(aXl + (unsigned)st * STAGE, po);
            gather_x(po, Pf);
            st = (st + 1 == S) ? 0 : st + 1;
            sti = (sti + 1 == S) ? 0 : sti + 1;
        }
        for (int k = N - 2; k >= 0; --k) {
            issue_bwd(k - D, sti);
            cp_wait(D);
            T q[RX], r[RU], Rf[NU];
            lds_piece<T, RX, SX>(aXl + (unsigned)st * STAGE, q);
            lds_piece<T, RU, SU>(aUl + (unsigned)st * STAGE, r);
            gather_u(r, Rf);
            // d_k = Quu_inv ((B^T p_{k+1} + r_k) + BPf)
            T s_[RU], Sf[NU], acc1[RX + RU], kr[RX], dq[RU];
            dots<FAST>(mS1b, Pf, acc1);  // [AmBKt p_{k+1} ; B^T p_{k+1}]
#pragma unroll
            for (int b = 0; b < RU; ++b) s_[b] = (acc1[RX + b] + r[b]) + vBPf[b];
            gather_u(s_, Sf);
            // p_k = ((q_k + AmBKt p_{k+1}) - Kinf^T r_k) + APf
            dots<FAST>(mKt, Rf, kr);
#pragma unroll
            for (int a = 0; a < RX; ++a) po[a] = ((q[a] + acc1[a]) - kr[a]) + vAPf[a];
            if (k > 0) gather_x(po, Pf);  // p_0 itself is never used
            dots<FAST>(mQuu, Sf, dq);
            if (uvl) stg_piece<T, RU, CU>(gu + (int64_t)k * rec + LY.d, dq);
            st = (st + 1 == S) ? 0 : st + 1;
            sti = (sti + 1 == S) ? 0 : sti + 1;
        }
    };

    // ---- cooperative (all 32 lanes) load of instance `ib` into slot `s`: the slot's records are initialised from the
    // warm-start state (zeros when cold), including the first iteration's linear cost (admm.cpp:262-304 on the state
    // as solve() finds it, cone / hyperplane slacks = previous rollout, admm.cpp:352-376) ----
    auto load_slot = [&](int s, int64_t ib) {
        const int64_t ox = ib * (int64_t)N * NX, ou = ib * (int64_t)(N - 1) * NU;
        const T *xrefb = P.Xref + (P.xref_pi ? ox : 0);
        const T *urefb = has_uref ? P.Uref + (P.uref_pi ? ou : 0) : nullptr;
        const T *const sgf[3] = {P.s_gc, P.s_gl, P.s_gl_tv};
        const T *const syf[3] = {P.s_yc, P.s_yl, P.s_yl_tv};
        for (int e = lane; e < N * NX; e += 32) {
            const int k = e / NX, i = e - k * NX;
            const T vnew_in = (!cold && P.s_vnew) ? P.s_vnew[ox + e] : T(0);
            const T g_in = (!cold && P.s_g) ? P.s_g[ox + e] : T(0);
            const T v_in = (!cold && P.s_v) ? P.s_v[ox + e] : T(0);
            T acc;
            if (k < N - 1) {
                acc = -(__ldg(xrefb + e) * __ldg(gmat + OFF_QD + i));
            } else {  // -(Pinf^T xref_{N-1})(i), m ascending
                const T *xl = xrefb + (int64_t)(N - 1) * NX;
                T sacc = __ldg(xl) * __ldg(gmat + OFF_PINF + NX * i);
                for (int m = 1; m < NX; ++m) sacc = mac<FAST>(sacc, __ldg(xl + m), __ldg(gmat + OFF_PINF + m + NX * i));
                acc = -sacc;
            }
            acc = nmac<FAST>(acc, rho, vnew_in - g_in);
            T *r_ = wsw + (int64_t)k * rec + s * NX + i;
            r_[LY.vnew] = v_in;  // the slot of the box slack holds work->v until the first forward sweep rewrites it
            r_[LY.g] = g_in;
            if constexpr (EXT) {
                const T xin = (k == 0) ? __ldg(P.x0 + ib * NX + i) : ((!cold && P.s_x) ? P.s_x[ox + e] : T(0));
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (fx[f]) {
                        const T gf_in = (!cold && sgf[f]) ? sgf[f][ox + e] : T(0);
                        acc = nmac<FAST>(acc, rho, xin - gf_in);
                        r_[LY.gf[f]] = gf_in;
                        if (LY.vf[f] >= 0) r_[LY.vf[f]] = xin;
                    }
                }
            }
            r_[LY.q] = acc;
        }
        for (int e = lane; e < (N - 1) * NU; e += 32) {
            const int k = e / NU, j = e - k * NU;
            const T znew_in = (!cold && P.s_znew) ? P.s_znew[ou + e] : T(0);
            const T y_in = (!cold && P.s_y) ? P.s_y[ou + e] : T(0);
            const T z_in = (!cold && P.s_z) ? P.s_z[ou + e] : T(0);
            const T ur = has_uref ? __ldg(urefb + e) : T(0);
            T acc = nmac<FAST>(-(ur * __ldg(gmat + OFF_RD + j)), rho, znew_in - y_in);
            T *r_ = wsw + (int64_t)k * rec + s * NU + j;
            r_[LY.znew] = z_in;
            r_[LY.y] = y_in;
            if constexpr (EXT) {
                const T uin = (!cold && P.s_u) ? P.s_u[ou + e] : T(0);
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (fu[f]) {
                        const T yf_in = (!cold && syf[f]) ? syf[f][ou + e] : T(0);
                        acc = nmac<FAST>(acc, rho, uin - yf_in);
                        r_[LY.yf[f]] = yf_in;
                        if (LY.zf[f] >= 0) r_[LY.zf[f]] = uin;
                    }
                }
            }
            r_[LY.r] = acc;
        }
        if (slot == s) {
            inst = ib;
            busy = true;
            it = 0;
            solved = 0;
            res_px = res_dx = res_pu = res_du = T(0);
            xrefp = xrefb + l * RX;
            urefp = has_uref ? urefb + l * RU : P.Xref;
            const T *xl = xrefb + (int64_t)(N - 1) * NX;
#pragma unroll
            for (int a = 0; a < RX; ++a) {
                const int ii = xvl ? l * RX + a : 0;
                x0o[a] = xvl ? __ldg(P.x0 + ib * NX + ii) : T(0);
                T sacc = __ldg(xl) * __ldg(gmat + OFF_PINF + NX * ii);
                for (int m = 1; m < NX; ++m) sacc = mac<FAST>(sacc, __ldg(xl + m), __ldg(gmat + OFF_PINF + m + NX * ii));
                pterm[a] = xvl ? -sacc : T(0);
            }
        }
        __syncwarp();  // the records were written by all lanes; their owners read them from here on
    };

    // ---- cooperative write-back of slot `s` (instance `ib`) ----
    auto store_slot = [&](int s, int64_t ib) {
        const int s_solved = __shfl_sync(0xffffffffu, solved, s * L);
        const int s_it = __shfl_sync(0xffffffffu, it, s * L);
        if (slot == s && l == 0) {
            if (P.iter) P.iter[ib] = it;
            if (P.solved) P.solved[ib] = solved;
            if (P.residuals) {
                T *r = P.residuals + 4 * ib;
                r[0] = res_px; r[1] = res_dx; r[2] = res_pu; r[3] = res_du;
            }
        }
        __syncwarp();  // owner lanes wrote the records; every lane reads them below
        const int64_t ox = ib * (int64_t)N * NX, ou = ib * (int64_t)(N - 1) * NU;
        const bool ran = s_it > 0;
        T *const ovf[3] = {P.s_vcnew, P.s_vlnew, P.s_vlnew_tv};
        T *const ogf[3] = {P.s_gc, P.s_gl, P.s_gl_tv};
        T *const ozf[3] = {P.s_zcnew, P.s_zlnew, P.s_zlnew_tv};
        T *const oyf[3] = {P.s_yc, P.s_yl, P.s_yl_tv};
        for (int e = lane; e < N * NX; e += 32) {
            const int k = e / NX, i = e - k * NX;
            const T *r_ = wsw + (int64_t)k * rec + s * NX + i;
            // solution->x = vnew (admm.cpp:436,452); no iteration (max_iter <= 0): the state as it came in
            const T v = ran ? r_[LY.vnew] : ((!cold && P.s_vnew) ? P.s_vnew[ox + e] : T(0));
            if (P.sol_x) P.sol_x[ox + e] = v;
            if (P.s_vnew) P.s_vnew[ox + e] = v;
            if (P.s_g) P.s_g[ox + e] = r_[LY.g];
            // work->v: previous vnew when the solve converged (the return at admm.cpp:441 precedes :445), else = vnew
            if (P.s_v && ran) P.s_v[ox + e] = s_solved ? r_[LY.vprev] : v;
            else if (P.s_v && cold) P.s_v[ox + e] = T(0);
            if constexpr (EXT) {
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (fx[f]) {
                        if (ovf[f]) ovf[f][ox + e] = r_[LY.vf[f]];
                        if (ogf[f]) ogf[f][ox + e] = r_[LY.gf[f]];
                    }
                }
            }
        }
        for (int e = lane; e < (N - 1) * NU; e += 32) {
            const int k = e / NU, j = e - k * NU;
            const T *r_ = wsw + (int64_t)k * rec + s * NU + j;
            const T z = ran ? r_[LY.znew] : ((!cold && P.s_znew) ? P.s_znew[ou + e] : T(0));
            if (P.sol_u) P.sol_u[ou + e] = z;
            if (P.s_znew) P.s_znew[ou + e] = z;
            if (P.s_y) P.s_y[ou + e] = r_[LY.y];
            if (P.s_z && ran) P.s_z[ou + e] = s_solved ? r_[LY.zprev] : z;
            else if (P.s_z && cold) P.s_z[ou + e] = T(0);
            if constexpr (EXT) {
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (fu[f]) {
                        if (ozf[f]) ozf[f][ou + e] = r_[LY.zf[f]];
                        if (oyf[f]) oyf[f][ou + e] = r_[LY.yf[f]];
                    }
                }
            }
        }
        // work->x / work->u (and u0 = work->u.col(0)): replay of the last rollout from d and x0, bit-identical to the last
        // forward sweep.  Every lane executes the arithmetic (the gathers are warp-wide); the lanes of slot s store.
        if (P.s_x || P.s_u || P.u0) {
            __syncwarp();
            const bool mine = slot == s;
            T xo[RX], Xf[NX];
#pragma unroll
            for (int a = 0; a < RX; ++a) xo[a] = x0o[a];
            gather_x(xo, Xf);
            const int kend = (P.s_x || P.s_u) ? N : 1;
            for (int k = 0; k < kend; ++k) {
                if (P.s_x && mine && xvl) {
#pragma unroll
                    for (int a = 0; a < RX; ++a) {
                        if (ran || k == 0) P.s_x[ox + (int64_t)k * NX + l * RX + a] = xo[a];
                        else if (cold) P.s_x[ox + (int64_t)k * NX + l * RX + a] = T(0);
                    }
                }
                if (k < N - 1) {
                    T u[RU], Uf[NU], t1[RX + RU], bu_[RX], dk[RU];
#pragma unroll
                    for (int b = 0; b < RU; ++b) dk[b] = (ran && uvl) ? gu[(int64_t)k * rec + LY.d + b] : T(0);
                    dots<FAST>(mS1f, Xf, t1);
#pragma unroll
                    for (int b = 0; b < RU; ++b) u[b] = (-t1[RX + b]) - dk[b];
                    if (mine && uvl) {
#pragma unroll
                        for (int b = 0; b < RU; ++b) {
                            const int64_t o = ou + (int64_t)k * NU + l * RU + b;
                            if (P.s_u) {
                                if (ran) P.s_u[o] = u[b];
                                else if (cold) P.s_u[o] = T(0);
                            }
                            if (P.u0 && k == 0) P.u0[ib * NU + l * RU + b] = ran ? u[b] : ((!cold && P.s_u) ? P.s_u[o] : T(0));
                        }
                    }
                    if (k + 1 < kend) {
                        gather_u(u, Uf);
                        dots<FAST>(mB, Uf, bu_);
#pragma unroll
                        for (int a = 0; a < RX; ++a) xo[a] = (t1[a] + bu_[a]) + vf[a];
                        gather_x(xo, Xf);
                    }
                }
            }
        }
        __syncwarp();
    };

    // ---- persistent loop (same protocol as the on-chip kernel): retire / refill slots, then iterate until some
    // slot terminates; the iteration loop has warp-uniform control flow only ----
    for (;;) {
        const bool fin = busy && (solved || it >= P.max_iter);
        const unsigned todo = __ballot_sync(0xffffffffu, (fin || (!busy && want)) && l == 0);
        for (unsigned m = todo; m; m &= m - 1) {
            const int s = (__ffs(m) - 1) / L;
            const int64_t ib_old = __shfl_sync(0xffffffffu, inst, s * L);
            const int was_busy = __shfl_sync(0xffffffffu, (int)busy, s * L);
            unsigned long long nxt = 0;
            if (lane == 0) nxt = atomicAdd(queue, 1ULL);
            if (was_busy) store_slot(s, ib_old);
            nxt = __shfl_sync(0xffffffffu, nxt, 0);
            if ((int64_t)nxt < P.B) {
                load_slot(s, (int64_t)nxt);
            } else if (slot == s) {
                busy = false;
                want = false;
            }
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (__any_sync(0xffffffffu, busy && it >= P.max_iter)) continue;  // max_iter <= 0: retire without iterating
        __syncwarp();
        do {
            backward();
            __syncwarp();
            T rpx = T(0), rdx = T(0), rpu = T(0), rdu = T(0);
            forward(rpx, rdx, rpu, rdu);
            __syncwarp();
            // termination_condition (admm.cpp:310-328), per instance
            rpx = group_max<T, L>(rpx);
            rdx = group_max<T, L>(rdx);
            rpu = group_max<T, L>(rpu);
            rdu = group_max<T, L>(rdu);
            if (busy) {
                it += 1;
                if (it % P.check_termination == 0) {
                    res_px = rpx;
                    res_dx = rdx * rho;
                    res_pu = rpu;
                    res_du = rdu * rho;
                    if (res_px < P.pri_tol && res_pu < P.pri_tol && res_dx < P.dua_tol && res_du < P.dua_tol) solved = 1;
                }
            }
        } while (!__any_sync(0xffffffffu, busy && (solved || it >= P.max_iter)));
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: record layout, resident-slot plan, launch
// ---------------------------------------------------------------------------------------------------------
struct GpsPlan {
    int L = 0, warps = 0, ctas = 0, dist = 0;
    size_t smem = 0, ws_bytes = 0;
    GpsLayout ly;
};

inline int gps_env_int(const char *name, int dflt) {
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

template <typename T, int NX, int NU, int L, bool EXT>
inline GpsPlan gps_plan_L(const LaunchDesc &d) {
    using Cfg = GpsCfg<NX, NU, L, (int)sizeof(T), EXT>;
    GpsPlan p;
    GpsLayout &ly = p.ly;
    const int ES = (int)sizeof(T), IPW = Cfg::IPW;
    int off = 0;
    auto take = [&](bool present, int rows) {
        if (!present) return -1;
        const int o = off;
        off += (IPW * rows * ES + 15) / 16 * 16 / ES;
        return o;
    };
    const tinympc_state_t &s = d.io.state;
    const bool fx[3] = {EXT && d.soc_x, EXT && d.lin_x, EXT && d.tvl_x}, fu[3] = {EXT && d.soc_u, EXT && d.lin_u, EXT && d.tvl_u};
    const void *ovf[3] = {s.vcnew, s.vlnew, s.vlnew_tv}, *ozf[3] = {s.zcnew, s.zlnew, s.zlnew_tv};
    ly.d = take(true, NU);
    ly.vnew = take(true, NX);
    ly.g = take(true, NX);
    for (int f = 0; f < 3; ++f) ly.gf[f] = take(fx[f], NX);
    ly.znew = take(true, NU);
    ly.y = take(true, NU);
    for (int f = 0; f < 3; ++f) ly.yf[f] = take(fu[f], NU);
    ly.q = take(true, NX);
    ly.r = take(true, NU);
    const bool keep_v = s.v || s.z;
    ly.vprev = take(keep_v, NX);
    ly.zprev = take(keep_v, NU);
    for (int f = 0; f < 3; ++f) ly.vf[f] = take(fx[f] && ovf[f], NX);
    for (int f = 0; f < 3; ++f) ly.zf[f] = take(fu[f] && ozf[f], NU);
    ly.rec = off;
    int dist = gps_env_int("TINYMPC_GPS_DIST", 2);
    dist = std::max(1, std::min(3, dist));
    const int max_smem = d.max_smem_optin - 64;
    const size_t blob = ((size_t)(3 * NX * NX + 2 * NX * NU + NU * NU + 4 * NX + 2 * NU) * sizeof(T) + 15) / 16 * 16 + 64;
    while (dist > 1 && Cfg::warp_bytes(dist + 1) * 4 > (size_t)max_smem) --dist;  // keep at least four warps per CTA
    ly.dist = dist;
    const size_t per_warp = Cfg::warp_bytes(dist + 1);
    int maxw = (int)std::min<size_t>(GPS_MAX_WARPS, (size_t)max_smem / per_warp);
    maxw = std::max(1, std::min(maxw, std::max(1, gps_env_int("TINYMPC_GPS_WARPS", GPS_MAX_WARPS))));
    if (per_warp > (size_t)max_smem || blob > (size_t)max_smem) return p;
    // balance the waves: with `waves` passes over the resident slots, use just enough warps per SM to hold B / waves
    const int64_t groups = (d.io.B + IPW - 1) / IPW;
    const int64_t cap = (int64_t)d.sm_count * maxw;
    const int64_t waves = std::max<int64_t>(1, (groups + cap - 1) / cap);
    int warps = (int)std::min<int64_t>(maxw, std::max<int64_t>(1, (groups + waves * d.sm_count - 1) / (waves * d.sm_count)));
    p.L = L;
    p.warps = warps;
    p.ctas = (int)std::max<int64_t>(1, std::min<int64_t>(d.sm_count, (groups + warps - 1) / warps));
    p.dist = dist;
    p.smem = std::max(per_warp * (size_t)warps, blob);
    p.ws_bytes = (size_t)p.ctas * warps * d.N * ly.rec * sizeof(T);
    return p;
}

template <typename T, int NX, int NU, bool FAST, bool EXT>
int launch_gps(LaunchDesc *d, const KParams<T, NX, NU> &P0) {
    constexpr int L = gps_pick_L<T, NX, NU>();
    if constexpr (L == 0) {
        return TINYMPC_ERR_UNSUPPORTED;
    } else {
        const GpsPlan plan = gps_plan_L<T, NX, NU, L, EXT>(*d);
        if (plan.L == 0 || !d->gmat || !d->work_queue) return TINYMPC_ERR_UNSUPPORTED;
        d->out_ws_need = plan.ws_bytes;
        if (!d->gps_ws || d->gps_ws_bytes < plan.ws_bytes) return TM_ERR_WORKSPACE;
        KParams<T, NX, NU> P = P0;
        P.gps = plan.ly;
        P.gps_ws = (T *)d->gps_ws;
        auto kern = gps_solve_kernel<T, NX, NU, L, FAST, EXT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) != cudaSuccess) return TINYMPC_ERR_CUDA;
        kern<<<plan.ctas, plan.warps * 32, plan.smem, d->stream>>>(P, (const T *)d->gmat, (unsigned long long *)d->work_queue);
        d->out_threads = plan.warps * 32;
        d->out_ctas = plan.ctas;
        d->out_smem = (int)plan.smem;
        d->out_lanes_per_instance = L;
        d->out_instances_per_cta = plan.warps * (32 / L);
        d->out_tmem_cols = 0;
        return cudaGetLastError() == cudaSuccess ? TINYMPC_OK : TINYMPC_ERR_CUDA;
    }
}

}  // namespace tmpc

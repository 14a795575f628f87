// kernels_inst.cu — compiled once per supported (nx, nu) with -DTM_NX=.. -DTM_NU=.. (see Makefile).
// Instantiates the TPI (and, when it fits, GPI) kernels for float/double x strict/fast x box/extended
// and exports one type-erased launcher  tm_dim_entry_<nx>_<nu>().
#include <algorithm>
#include <cstring>
#include <limits>

#include "launch.h"
#include "tpi_kernel.cuh"
#ifdef TM_WITH_GPI
#include "gpi_kernel.cuh"
#endif
#include "precompute_kernel.cuh"

#ifndef TM_NX
#error "compile with -DTM_NX=<nx> -DTM_NU=<nu>"
#endif

namespace tmpc {
namespace {

template <typename T>
void fill_params(KParams<T, TM_NX, TM_NU> &P, const LaunchDesc &d) {
    constexpr int NX = TM_NX, NU = TM_NU;
    std::memset(&P, 0, sizeof(P));
    std::memcpy(P.A, d.A, sizeof(T) * NX * NX);
    std::memcpy(P.Bm, d.Bm, sizeof(T) * NX * NU);
    std::memcpy(P.f, d.f, sizeof(T) * NX);
    std::memcpy(P.Qd, d.Qd, sizeof(T) * NX);
    std::memcpy(P.Rd, d.Rd, sizeof(T) * NU);
    std::memcpy(P.Kinf, d.Kinf, sizeof(T) * NU * NX);
    std::memcpy(P.Pinf, d.Pinf, sizeof(T) * NX * NX);
    std::memcpy(P.Quu, d.Quu, sizeof(T) * NU * NU);
    std::memcpy(P.AmBKt, d.AmBKt, sizeof(T) * NX * NX);
    std::memcpy(P.APf, d.APf, sizeof(T) * NX);
    std::memcpy(P.BPf, d.BPf, sizeof(T) * NU);
    P.rho = (T)d.rho;
    P.pri_tol = (T)d.pri_tol;
    P.dua_tol = (T)d.dua_tol;
    P.N = d.N;
    P.max_iter = d.max_iter;
    P.check_termination = d.check_termination;
    P.en_state_bound = d.en_state_bound;
    P.en_input_bound = d.en_input_bound;
    P.soc_x = d.soc_x; P.soc_u = d.soc_u; P.ncx = d.ncx; P.ncu = d.ncu;
    P.lin_x = d.lin_x; P.lin_u = d.lin_u; P.nlx = d.nlx; P.nlu = d.nlu;
    P.tvl_x = d.tvl_x; P.tvl_u = d.tvl_u; P.ntvx = d.ntvx; P.ntvu = d.ntvu;
    for (int c = 0; c < MAX_CONES; ++c) {
        P.cone_x_start[c] = d.cone_x_start[c];
        P.cone_u_start[c] = d.cone_u_start[c];
        P.cone_x_mu[c] = (T)d.cone_x_mu[c];
        P.cone_u_mu[c] = (T)d.cone_u_mu[c];
    }
    const tinympc_batch_t &io = d.io;
    P.B = io.B;
    P.Bpad = d.Bpad;
    P.cold = io.cold_start;
    P.bounds_tv = d.bounds_tv;
    {
        const T inf = std::numeric_limits<T>::infinity();
        for (int i = 0; i < NX; ++i) {
            P.xlo[i] = (d.en_state_bound && d.h_xlo) ? ((const T *)d.h_xlo)[i] : -inf;
            P.xhi[i] = (d.en_state_bound && d.h_xhi) ? ((const T *)d.h_xhi)[i] : inf;
        }
        for (int j = 0; j < NU; ++j) {
            P.ulo[j] = (d.en_input_bound && d.h_ulo) ? ((const T *)d.h_ulo)[j] : -inf;
            P.uhi[j] = (d.en_input_bound && d.h_uhi) ? ((const T *)d.h_uhi)[j] : inf;
        }
    }
    P.Pinf_g = d.gmat ? (const T *)d.gmat + (NX * NX + NX * NU + NX + NX + NU + NU * NX) : nullptr;
    P.xref_pi = io.xref_per_instance;
    P.uref_pi = io.uref_per_instance;
    P.x0 = (const T *)io.x0; P.Xref = (const T *)io.Xref; P.Uref = (const T *)io.Uref;
    P.x_min = (const T *)d.x_min; P.x_max = (const T *)d.x_max; P.u_min = (const T *)d.u_min; P.u_max = (const T *)d.u_max;
    P.Alin_x = (const T *)d.Alin_x; P.blin_x = (const T *)d.blin_x; P.Alin_u = (const T *)d.Alin_u; P.blin_u = (const T *)d.blin_u;
    P.tv_Alin_x = (const T *)d.tv_Alin_x; P.tv_blin_x = (const T *)d.tv_blin_x;
    P.tv_Alin_u = (const T *)d.tv_Alin_u; P.tv_blin_u = (const T *)d.tv_blin_u;
    const tinympc_state_t &s = io.state;
    P.s_x = (T *)s.x; P.s_u = (T *)s.u; P.s_v = (T *)s.v; P.s_z = (T *)s.z;
    P.s_vnew = (T *)s.vnew; P.s_znew = (T *)s.znew; P.s_g = (T *)s.g; P.s_y = (T *)s.y;
    P.s_vcnew = (T *)s.vcnew; P.s_zcnew = (T *)s.zcnew; P.s_gc = (T *)s.gc; P.s_yc = (T *)s.yc;
    P.s_vlnew = (T *)s.vlnew; P.s_zlnew = (T *)s.zlnew; P.s_gl = (T *)s.gl; P.s_yl = (T *)s.yl;
    P.s_vlnew_tv = (T *)s.vlnew_tv; P.s_zlnew_tv = (T *)s.zlnew_tv; P.s_gl_tv = (T *)s.gl_tv; P.s_yl_tv = (T *)s.yl_tv;
    P.sol_x = (T *)io.sol_x; P.sol_u = (T *)io.sol_u;
    P.iter = io.iter; P.solved = io.solved; P.residuals = (T *)io.residuals;
    P.u0 = (T *)io.u0;
    P.models = (const T *)io.models;
    P.gpi_vscratch = (T *)d.gpi_vscratch;
    P.w_v[0] = d.w_v[0]; P.w_v[1] = d.w_v[1]; P.w_z[0] = d.w_z[0]; P.w_z[1] = d.w_z[1];
    P.w_g = d.w_g; P.w_y = d.w_y; P.w_d = d.w_d;
    P.w_vc = d.w_vc; P.w_zc = d.w_zc; P.w_gc = d.w_gc; P.w_yc = d.w_yc;
    P.w_vl = d.w_vl; P.w_zl = d.w_zl; P.w_gl = d.w_gl; P.w_yl = d.w_yl;
    P.w_vlt = d.w_vlt; P.w_zlt = d.w_zlt; P.w_glt = d.w_glt; P.w_ylt = d.w_ylt;
}

template <typename T, bool FAST, bool EXT>
int launch_tpi(LaunchDesc *d) {
    KParams<T, TM_NX, TM_NU> P;
    fill_params<T>(P, *d);
    const int threads = TPI_THREADS;
    const int64_t blocks = (d->io.B + threads - 1) / threads;
    if (blocks <= 0) return TINYMPC_OK;
    tpi_solve_kernel<T, TM_NX, TM_NU, FAST, EXT><<<(unsigned)blocks, threads, 0, d->stream>>>(P);
    d->out_threads = threads;
    d->out_ctas = (int)blocks;
    d->out_smem = 0;
    d->out_lanes_per_instance = 1;
    d->out_instances_per_cta = threads;
    d->out_tmem_cols = 0;
    return cudaGetLastError() == cudaSuccess ? TINYMPC_OK : TINYMPC_ERR_CUDA;
}

#ifdef TM_WITH_GPI
template <typename T, int NX, int NU, bool FAST>
int launch_gpi(LaunchDesc *d) {
    const GpiPlan plan = gpi_plan<T, NX, NU>(d->N, d->max_smem_optin - 64);
    if (plan.L == 0 || !d->gmat || !d->work_queue) return TINYMPC_ERR_UNSUPPORTED;
    KParams<T, NX, NU> P;
    fill_params<T>(P, *d);
    const T *gmat = (const T *)d->gmat;
    const bool het = d->io.models != nullptr;  // heterogeneous batch: per-instance model blobs
#define TM_GPI_CASE(LL, HH, TT) \
    if (plan.L == LL && het == HH && plan.tm == TT) return launch_gpi_L<T, NX, NU, LL, FAST, HH, TT>(d, plan, P, gmat);
#define TM_GPI_L(LL) TM_GPI_CASE(LL, false, false) TM_GPI_CASE(LL, true, false) TM_GPI_CASE(LL, false, true) TM_GPI_CASE(LL, true, true)
    TM_GPI_L(4)
    TM_GPI_L(8)
#ifdef TM_GPI_L16
    TM_GPI_L(16)
#else
    if constexpr (sizeof(T) == 8) {  // fp64 needs L = 16 from nx = 12 on (see gpi_plan)
        TM_GPI_CASE(16, false, false)
        TM_GPI_CASE(16, true, false)
    }
#endif
#undef TM_GPI_L
#undef TM_GPI_CASE
    return TINYMPC_ERR_UNSUPPORTED;
}
#endif

template <typename T>
int launch_T(LaunchDesc *d) {
#ifdef TM_WITH_GPI
    if (d->family == TINYMPC_KERNEL_GPI) {
        if (d->ext) return TINYMPC_ERR_UNSUPPORTED;
        return d->fast ? launch_gpi<T, TM_NX, TM_NU, true>(d) : launch_gpi<T, TM_NX, TM_NU, false>(d);
    }
#else
    if (d->family == TINYMPC_KERNEL_GPI) return TINYMPC_ERR_UNSUPPORTED;
#endif
    if (d->io.models) return TINYMPC_ERR_UNSUPPORTED;  // per-instance models live in the GPI kernel's per-lane registers only
    if (d->ext) return d->fast ? launch_tpi<T, true, true>(d) : launch_tpi<T, false, true>(d);
    return d->fast ? launch_tpi<T, true, false>(d) : launch_tpi<T, false, false>(d);
}

int launch(LaunchDesc *d) {
    if (d->dtype == TINYMPC_F32) return launch_T<float>(d);
    if (d->dtype == TINYMPC_F64) return launch_T<double>(d);
    return TINYMPC_ERR_ARG;
}

int gpi_fit(int dtype, int N, int max_smem_optin) {
#ifdef TM_WITH_GPI
    if (dtype == TINYMPC_F32) return gpi_fit_T<float, TM_NX, TM_NU>(N, max_smem_optin - 64);
    if (dtype == TINYMPC_F64) return gpi_fit_T<double, TM_NX, TM_NU>(N, max_smem_optin - 64);
#endif
    (void)dtype; (void)N; (void)max_smem_optin;
    return 0;
}

int gpi_ipc(int dtype, int N, int max_smem_optin) {
#ifdef TM_WITH_GPI
    if (dtype == TINYMPC_F32) {
        const GpiPlan p = gpi_plan<float, TM_NX, TM_NU>(N, max_smem_optin - 64);
        return p.L ? ((p.warps << 16) | (p.warps * (32 / p.L))) : 0;
    }
    if (dtype == TINYMPC_F64) {
        const GpiPlan p = gpi_plan<double, TM_NX, TM_NU>(N, max_smem_optin - 64);
        return p.L ? ((p.warps << 16) | (p.warps * (32 / p.L))) : 0;
    }
#endif
    (void)dtype; (void)N; (void)max_smem_optin;
    return 0;
}

int precompute_batch(int dtype, int64_t B, const void *A, const void *Bm, const void *f, const void *Qdiag, const void *Rdiag,
                     const void *rho, void *models_out, int32_t *sweeps_out, int sm_count, cudaStream_t stream) {
    if (dtype == TINYMPC_F32)
        return launch_precompute_T<float, TM_NX, TM_NU>(B, A, Bm, f, Qdiag, Rdiag, rho, models_out, sweeps_out, sm_count, stream);
    if (dtype == TINYMPC_F64)
        return launch_precompute_T<double, TM_NX, TM_NU>(B, A, Bm, f, Qdiag, Rdiag, rho, models_out, sweeps_out, sm_count, stream);
    return TINYMPC_ERR_ARG;
}

}  // namespace
}  // namespace tmpc

#define TM_CAT2(a, b, c) a##b##_##c
#define TM_CAT(a, b, c) TM_CAT2(a, b, c)
extern "C" const tmpc::DimEntry *TM_CAT(tm_dim_entry_, TM_NX, TM_NU)() {
    static const tmpc::DimEntry e = {TM_NX, TM_NU, &tmpc::launch, &tmpc::gpi_fit, &tmpc::gpi_ipc, &tmpc::precompute_batch};
    return &e;
}

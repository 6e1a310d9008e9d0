// swr.cu -- ns dispatch + launch of the sliding-window register kernels (3-D, float, ns <= 8).
#include <algorithm>

#include "rt2_kernels.cuh"
#include "swr2_kernels.cuh"

// 3-D: spreader k_swr2_spread (swr2_kernels.cuh: absolute ring slots), interpolator k_swr_interp
// (swr_kernels.cuh: phase chain + cp.async plane staging; an absolute-slot version was slower,
// DESIGN.md 4.3).

namespace b2n {

template <int NS> static void swr_bins_ns(int ns, int *bin) {
  if (ns == NS) {
    bin[0] = SwrCfg<NS>::BX;
    bin[1] = SwrCfg<NS>::BY;
    bin[2] = SwrCfg<NS>::BZ;
    return;
  }
  if constexpr (NS < 8) swr_bins_ns<NS + 1>(ns, bin);
}
// bins (in anchor cells) of the register kernels for kernel width ns: sliding window (3-D) or
// register tile (2-D, Rt2Cfg: (17 - ns)^2)
void swr_bins(int dim, int ns, int *bin, bool stacked2) {
  if (dim == 2) {
    bin[0] = stacked2 ? 9 - ns : 17 - ns;   // Rt2sCfg (8 x 12 window) : Rt2Cfg (16 x 16)
    bin[1] = stacked2 ? 4 * RT2S_S + 1 - ns : 17 - ns;
    bin[2] = 1;
    return;
  }
  swr_bins_ns<2>(ns, bin);
}

static void swr_fill(Plan<float> &p, SwrArgs &a) {
  a.rec = p.pts.rec;
  a.bin_start = p.pts.bin_start;
  a.sp_off = p.pts.sp_off;
  a.sp_bin = p.pts.sp_bin;
  a.M = p.pts.M;
  a.nftot = p.nftot;
  for (int d = 0; d < 3; d++) {
    a.nf[d] = (int)p.nf[d];
    a.bin[d] = p.bin[d];
    a.nbin[d] = p.nbin[d];
  }
  a.nbins = p.nbins;
  a.maxsub = p.maxsub;
}

template <int NS> struct SwrDispatch {
  static int spread(Plan<float> &p, const SwrArgs &a, int ntr) {
    if (p.ns == NS) {
      using C = Swr2Cfg<NS>;
      dim3 grid((unsigned)p.pts.sp_cap, (unsigned)ntr);
      if (a.scale) {  // type 3: strengths times the prephase
        B2N_CUDA_OK(cudaFuncSetAttribute(k_swr2_spread<NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SPREAD_SMEM));
        k_swr2_spread<NS, true><<<grid, 32, C::SPREAD_SMEM, p.stream>>>(a, p.tab);
      } else {
        B2N_CUDA_OK(cudaFuncSetAttribute(k_swr2_spread<NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SPREAD_SMEM));
        k_swr2_spread<NS, false><<<grid, 32, C::SPREAD_SMEM, p.stream>>>(a, p.tab);
      }
      B2N_LAUNCHED(1);
      B2N_LAUNCH_OK();
      return 0;
    }
    return SwrDispatch<NS + 1>::spread(p, a, ntr);
  }
  static int interp(Plan<float> &p, const SwrArgs &a, int ntr) {
    if (p.ns == NS) {
      using C = SwrCfg<NS>;
      const size_t smem = SwrInterpSmem<NS>::bytes();
      B2N_CUDA_OK(cudaFuncSetAttribute(k_swr_interp<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dim3 grid((unsigned)cdiv(p.pts.sp_cap, C::WARPS), (unsigned)ntr);
      k_swr_interp<NS><<<grid, 32 * C::WARPS, smem, p.stream>>>(a, p.tab);  B2N_LAUNCHED(1);
      B2N_LAUNCH_OK();
      return 0;
    }
    return SwrDispatch<NS + 1>::interp(p, a, ntr);
  }
};
template <> struct SwrDispatch<9> {
  static int spread(Plan<float> &, const SwrArgs &, int) { return B2N_ERR_METHOD_NOTVALID; }
  static int interp(Plan<float> &, const SwrArgs &, int) { return B2N_ERR_METHOD_NOTVALID; }
};

template <int NS> struct Rt2Dispatch {
  static int spread(Plan<float> &p, const SwrArgs &a, int ntr) {
    if (p.ns == NS) {
      using C = Rt2Cfg<NS>;
      if (ntr > 1 && !a.scale) {  // stacked transforms sharing the points: RT2_NT of them per warp pass
        constexpr int NT = 4;
        using R = Rt2NtCfg<NS, NT>;
        dim3 gridn((unsigned)cdiv(p.pts.sp_cap, C::WARPS), (unsigned)cdiv(ntr, NT));
        B2N_CUDA_OK(cudaFuncSetAttribute(k_rt2_spread_nt<NS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::smem()));
        k_rt2_spread_nt<NS, NT><<<gridn, 32 * C::WARPS, R::smem(), p.stream>>>(a, p.tab, ntr);  B2N_LAUNCHED(1);
        B2N_LAUNCH_OK();
        return 0;
      }
      dim3 grid((unsigned)cdiv(p.pts.sp_cap, C::WARPS), (unsigned)ntr);
      B2N_CUDA_OK(cudaFuncSetAttribute(k_rt2_spread<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::spread_smem()));
      k_rt2_spread<NS><<<grid, 32 * C::WARPS, C::spread_smem(), p.stream>>>(a, p.tab);  B2N_LAUNCHED(1);
      B2N_LAUNCH_OK();
      return 0;
    }
    return Rt2Dispatch<NS + 1>::spread(p, a, ntr);
  }
  static int interp(Plan<float> &p, const SwrArgs &a, int ntr) {
    if (p.ns == NS) {
      using C = Rt2Cfg<NS>;
      if (ntr > 1 && !a.scale) {  // stacked transforms sharing the points
        constexpr int NT = 4;
        using R = Rt2NtiCfg<NS, NT>;
        dim3 gridn((unsigned)cdiv(p.pts.sp_cap, C::WARPS), (unsigned)cdiv(ntr, NT));
        B2N_CUDA_OK(cudaFuncSetAttribute(k_rt2_interp_nt<NS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::smem()));
        k_rt2_interp_nt<NS, NT><<<gridn, 32 * C::WARPS, R::smem(), p.stream>>>(a, p.tab, ntr);  B2N_LAUNCHED(1);
        B2N_LAUNCH_OK();
        return 0;
      }
      dim3 grid((unsigned)cdiv(p.pts.sp_cap, C::WARPS), (unsigned)ntr);
      B2N_CUDA_OK(cudaFuncSetAttribute(k_rt2_interp<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::interp_smem()));
      k_rt2_interp<NS><<<grid, 32 * C::WARPS, C::interp_smem(), p.stream>>>(a, p.tab);  B2N_LAUNCHED(1);
      B2N_LAUNCH_OK();
      return 0;
    }
    return Rt2Dispatch<NS + 1>::interp(p, a, ntr);
  }
};
template <> struct Rt2Dispatch<9> {
  static int spread(Plan<float> &, const SwrArgs &, int) { return B2N_ERR_METHOD_NOTVALID; }
  static int interp(Plan<float> &, const SwrArgs &, int) { return B2N_ERR_METHOD_NOTVALID; }
};

template <int NS> static int rt2s_spread(Plan<float> &p, const SwrArgs &a, const float2 *c, int ntr) {
  if (p.ns == NS) {
    using C = Rt2sCfg<NS>;
    const int64_t M = p.pts.M;
    const int npass = (ntr + C::NT - 1) / C::NT, cstride = npass * C::NT;
    if (M * cstride > p.cap_cpack) {
      dev_free(p.cpack, p.stream);
      p.cpack = nullptr;
      p.cap_cpack = 0;
      if (int e = dev_alloc_t(&p.cpack, (size_t)M * cstride, p.stream)) return e;
      p.cap_cpack = M * cstride;
    }
    B2N_CUDA_OK(cudaFuncSetAttribute(k_rt2s_spread<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    // all ntr transforms in one launch: the kernel runs its passes of NT back to back
    k_pack_strengths<<<(unsigned)cdiv(M, 32), 256, 0, p.stream>>>(c, M, ntr, cstride, p.cpack);
    k_rt2s_spread<NS><<<(unsigned)p.pts.sp_cap, 32, C::SMEM, p.stream>>>(a, p.tab, p.cpack, ntr, cstride);
    B2N_LAUNCHED(2);
    B2N_LAUNCH_OK();
    return 0;
  }
  if constexpr (NS < 7) return rt2s_spread<NS + 1>(p, a, c, ntr);
  return B2N_ERR_METHOD_NOTVALID;
}

int spread_swr(Plan<float> &p, const float2 *c, const float2 *prescale, float2 *fw, int ntr) {
  if (p.pts.M == 0 || p.pts.sp_cap == 0) return 0;
  SwrArgs a;
  swr_fill(p, a);
  a.cin = c;
  a.cout = nullptr;
  a.scale = prescale;
  a.fw = fw;
  if (p.dim == 2 && p.stacked2) return rt2s_spread<2>(p, a, c, ntr);  // the bins are the narrow-window ones
  if (p.dim == 2) return Rt2Dispatch<2>::spread(p, a, ntr);
  return SwrDispatch<2>::spread(p, a, ntr);
}

int interp_swr(Plan<float> &p, float2 *c, const float2 *postscale, const float2 *fw, int ntr) {
  if (p.pts.M == 0 || p.pts.sp_cap == 0) return 0;
  SwrArgs a;
  swr_fill(p, a);
  a.cin = nullptr;
  a.cout = c;
  a.scale = postscale;
  a.fw = const_cast<float2 *>(fw);
  if (p.dim == 2) return Rt2Dispatch<2>::interp(p, a, ntr);
  return SwrDispatch<2>::interp(p, a, ntr);
}

}  // namespace b2n

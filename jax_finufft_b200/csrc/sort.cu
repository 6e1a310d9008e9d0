// sort.cu -- bin-sort of the non-uniform points (setpts for types 1/2, and both point sets of
// type 3).  Replaces calc_bin_size_noghost_* + thrust scans + calc_inverse_of_global_sort_index_*
// + calc_subprob_* + map_b_into_subprob_* (V/src/cuda/3d/spreadinterp3d.cuh:28-84,
// V/src/cuda/precision_independent.cu:37-91, V/src/cuda/3d/spread3d_wrapper.cu:418-506).
//
// Contract kept from the reference: bin id = binx + biny*nbinx + binz*nbinx*nbiny
// with bin_d = clamp(floor(fold_rescale(x_d)/binsize_d)); bins concatenated by exclusive scan;
// order inside a bin unspecified -- which leaves us free to ORDER the points of a bin by a
// sub-key (the fine-grid z cell) that the 3-D sliding-window kernels (swr_kernels.cuh) rely on.
//
// B200 design.  The output is an array of 16-byte records {x', y', z', idx} (x' = folded
// coordinate) in key order, so spread/interp stream one LDG.128 per point.  A single-pass
// scatter of 1e8 such records writes partial 32-byte sectors all over a 1.6 GB array and was
// measured at 13.3 GB DRAM read + 12.2 GB write (8x amplification, profiles/r01a_*).  Instead:
//   P0  k_key_hist    key histogram, warp-aggregated RED.ADD (no per-point output)
//       scan          key_start[K+1]
//   P1  k_partition   CTA-local split of a 4096-point chunk into <= 256 buckets of consecutive
//                     keys; each (CTA, bucket) run is reserved with one atomicAdd and written
//                     together, so DRAM sees (nearly) full sectors                [skipped when
//                     the whole record array fits in L2]
//   P2  k_place       bucket by bucket: pos = key_start[key] + (atomicSub(count[key]) - 1); the
//                     scattered 16-byte stores of a bucket land in a window of a few MB that
//                     stays in L2 until its sectors are complete
// No host readback, no stream sync.  HBM bytes per point (float, 3-D): 12 + (12+16) + (16+16)
// = 72 B algorithmic.
#include <mutex>

#include "swr_kernels.cuh"

namespace b2n {

// All plan buffers come from a PRIVATE stream-ordered pool per device (not the device's default
// pool, which other libraries in the process share): its release threshold is lifted so that
// memory is kept across synchronisations -- re-mapping gigabytes per call costs far more than the
// transforms themselves -- and b2n_cache_clear() / b2n_trim() hand it back explicitly.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pool[64] = {};
cudaMemPool_t plan_pool(int dev) {
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pool[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t mp = nullptr;
    if (cudaMemPoolCreate(&mp, &props) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
    g_pool[dev] = mp;
  }
  return g_pool[dev];
}
// bytes the pool of the current device holds from the driver (reserved) and has handed out (used)
void pool_usage(size_t *reserved, size_t *used) {
  *reserved = *used = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pool[dev]) return;
  unsigned long long r = 0, u = 0;
  cudaMemPoolGetAttribute(g_pool[dev], cudaMemPoolAttrReservedMemCurrent, &r);
  cudaMemPoolGetAttribute(g_pool[dev], cudaMemPoolAttrUsedMemCurrent, &u);
  *reserved = (size_t)r;
  *used = (size_t)u;
}
void pool_trim(size_t keep) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (g_pool[dev]) cudaMemPoolTrimTo(g_pool[dev], keep);
}

int dev_alloc(void **p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) bytes = 16;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t mp = plan_pool(dev);
  cudaError_t e = mp ? cudaMallocFromPoolAsync(p, bytes, mp, st) : cudaMallocAsync(p, bytes, st);
  if (e != cudaSuccess) {
    fprintf(stderr, "[b200nufft] device allocation of %zu bytes failed: %s\n", bytes, cudaGetErrorString(e));
    *p = nullptr;
    cudaGetLastError();
    return B2N_ERR_ALLOC;
  }
  return 0;
}
void dev_free(void *p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}

// Setpts cache: every kernel of the sort takes a device flag and returns at once when it is set
// (the flag is uniform over the grid and read before any barrier).  See binsort_points.
#define B2N_SKIP_IF(p) do { if ((p) != nullptr && *(p) != 0) return; } while (0)
#define B2N_GATE(g) do { if ((g).gate != nullptr && *(g).gate != (g).gate_val) return; } while (0)

// ------------------------------------------------------------------------------- scan (int32)
constexpr int SCAN_T = 256;
constexpr int SCAN_E = 8;  // elements per thread
constexpr int SCAN_CH = SCAN_T * SCAN_E;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *sm /*[32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) sm[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
    int t = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += n;
    }
    sm[lane] = t - s;  // exclusive warp offsets
    if (lane == 31) sm[32] = t;
  }
  __syncthreads();
  int res = sm[wid] + inc - v;
  *total = sm[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_blocks(const int32_t *__restrict__ in,
                                                          int32_t *__restrict__ out, int64_t n,
                                                          int32_t *__restrict__ bsum, const int *skip) {
  B2N_SKIP_IF(skip);
  __shared__ int sm[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CH + (int64_t)threadIdx.x * SCAN_E;
  int v[SCAN_E];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int off = block_exclusive_scan(s, &total, sm);
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    if (base + k < n) out[base + k] = off;
    off += v[k];
  }
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_bsum(int32_t *bsum, int nb, int32_t *grand, const int *skip) {
  B2N_SKIP_IF(skip);
  __shared__ int sm[33];
  int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    int i = b0 + threadIdx.x;
    int v = i < nb ? bsum[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total, sm);
    if (i < nb) bsum[i] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_add(int32_t *__restrict__ out, int64_t n,
                                                       const int32_t *__restrict__ bsum, const int *skip) {
  B2N_SKIP_IF(skip);
  const int64_t base = (int64_t)blockIdx.x * SCAN_CH + (int64_t)threadIdx.x * SCAN_E;
  const int add = bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_E; k++)
    if (base + k < n) out[base + k] += add;
}

// small inputs (1-D plans, the per-bin subproblem counts of small grids): one block, one launch --
// at M = 1e6 the sort is bound by launch gaps, not by bytes
constexpr int SCAN_SMALL = 1024 * SCAN_E;
__global__ void __launch_bounds__(1024) k_scan_small(const int32_t *__restrict__ in, int32_t *__restrict__ out, int n,
                                                     const int *skip) {
  B2N_SKIP_IF(skip);
  __shared__ int sm[33];
  const int base = threadIdx.x * SCAN_E;
  int v[SCAN_E];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    v[k] = base + k < n ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int off = block_exclusive_scan(s, &total, sm);
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    if (base + k < n) out[base + k] = off;
    off += v[k];
  }
  if (threadIdx.x == 0) out[n] = total;
}

// out[0..n]: out[i] = sum(in[0..i-1]); out[n] = total.  in/out may not alias.  `skip` (device flag,
// may be null): non-zero turns every launch into a no-op (setpts cache, binsort_points).
int exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t st, const int *skip) {
  if (n <= 0) {
    B2N_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
    return 0;
  }
  if (n <= SCAN_SMALL) {
    k_scan_small<<<1, 1024, 0, st>>>(in, out, (int)n, skip);  B2N_LAUNCHED(1);
    B2N_LAUNCH_OK();
    return 0;
  }
  const int nb = cdiv(n, SCAN_CH);
  int32_t *bsum = nullptr;
  if (int e = dev_alloc_t(&bsum, (size_t)nb + 1, st)) return e;
  k_scan_blocks<<<nb, SCAN_T, 0, st>>>(in, out, n, bsum, skip);  B2N_LAUNCHED(1);
  k_scan_bsum<<<1, 1024, 0, st>>>(bsum, nb, out + n, skip);  B2N_LAUNCHED(1);
  k_scan_add<<<nb, SCAN_T, 0, st>>>(out, n, bsum, skip);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  dev_free(bsum, st);
  return 0;
}

// ------------------------------------------------------------------------------- bin sort
struct SortGeom {
  int dim;
  int nf[3];
  int bin[3];
  int nbin[3];
  int nsub;    // 1, or bin[2]: sub-key = anchor z cell inside the bin (SWR kernels)
  int swr_ns;  // 0: reference bins (floor of the folded coordinate); else bins of ANCHOR cells
  unsigned magic[3];  // ceil(2^32 / bin[d]) (0 for bin 1): u / bin = umulhi(u, magic), u < 2^26
  const int *skip;    // setpts cache: non-zero = these points are already sorted in the plan (or null)
  // two-pass sort (binsort_points): the fast passes run while *gate == 0 (no bucket overflowed), the
  // three-pass fallback only once it is 1; null = unconditional
  const int *gate;
  int gate_val;
};

// u / g.bin[d] without an integer division (the three runtime divides were a third of the
// instructions of every sort pass): exact for 0 <= u < 2^32 / bin, bin <= 64  =>  u < 2^26
__device__ __forceinline__ int div_bin(const SortGeom &g, int d, int u) {
  return g.magic[d] ? (int)__umulhi((unsigned)u, g.magic[d]) : u;
}

// SWR geometry (3-D float): bins and sub-key are taken from the anchor cell of the ns-wide window
// (swr_kernels.cuh).  May shift a coordinate by -nf (periodic image), which is what gets stored.
__device__ __forceinline__ int swr_key(const SortGeom &g, float &xr, float &yr, float &zr) {
  const int ux = swr_anchor(xr, g.swr_ns, g.nf[0]);
  const int uy = swr_anchor(yr, g.swr_ns, g.nf[1]);
  if (g.dim == 2) return div_bin(g, 0, ux) + g.nbin[0] * div_bin(g, 1, uy);  // register-tile kernels: no sub-key
  const int uz = swr_anchor(zr, g.swr_ns, g.nf[2]);
  const int bz = div_bin(g, 2, uz);
  const int b = div_bin(g, 0, ux) + g.nbin[0] * (div_bin(g, 1, uy) + g.nbin[1] * bz);
  return b * g.nsub + (uz - bz * g.bin[2]);
}
__device__ __forceinline__ int swr_key(const SortGeom &, double &, double &, double &) { return 0; }

template <typename T>
__device__ __forceinline__ int point_key(const SortGeom &g, T &xr, T &yr, T &zr) {
  if (g.swr_ns) return swr_key(g, xr, yr, zr);
  int b = bin_of(xr, g.bin[0], g.nbin[0]);
  if (g.dim > 1) b += g.nbin[0] * bin_of(yr, g.bin[1], g.nbin[1]);
  if (g.dim > 2) {
    const int bz = bin_of(zr, g.bin[2], g.nbin[2]);
    b += g.nbin[0] * g.nbin[1] * bz;
    if (g.nsub > 1) {
      int sub = (int)zr - bz * g.bin[2];
      sub = sub < 0 ? 0 : (sub >= g.nsub ? g.nsub - 1 : sub);
      b = b * g.nsub + sub;
    }
  }
  return b;
}

template <typename T>
__device__ __forceinline__ void fold3(const SortGeom &g, int64_t i, const T *__restrict__ x,
                                      const T *__restrict__ y, const T *__restrict__ z, T &xr, T &yr,
                                      T &zr) {
  xr = fold_rescale(x[i], g.nf[0]);
  yr = g.dim > 1 ? fold_rescale(y[i], g.nf[1]) : T(0);
  zr = g.dim > 2 ? fold_rescale(z[i], g.nf[2]) : T(0);
}

template <typename T>
__device__ __forceinline__ PtRec<T> make_rec(T xr, T yr, T zr, int64_t i) {
  PtRec<T> r;
  r.x = xr; r.y = yr; r.z = zr;
  r.idx = (decltype(r.idx))i;
  return r;
}

// P0: key histogram.  One RED per distinct key per warp (__match_any_sync).  Keys that repeat
// inside a warp mean clustered input: a few hundred hot keys would then take ~1e8 / 20 REDs on the
// same few L2 sectors (measured 5.3 ms at C3-clustered vs 0.96 ms uniform, same-address REDs
// retire every ~50 ns).  Such keys are first summed per CTA in a small shared-memory hash table
// and reach L2 once per (CTA, key); `dupes` counts the merged lanes so that the placement pass can
// pick its clustered variant without a host round trip.  Uniform input never takes that path.
constexpr int HT_N = 1024;   // entries of the per-CTA hot-key table (power of two)
constexpr int HT_EMPTY = -1;
__device__ __forceinline__ int ht_slot(int key) { return (int)(((unsigned)key * 2654435761u) >> 22); }  // 10 bits

template <typename T>
__global__ void __launch_bounds__(256) k_key_hist(SortGeom g, int64_t M, const T *__restrict__ x,
                                                   const T *__restrict__ y, const T *__restrict__ z,
                                                   int32_t *__restrict__ hist, int32_t *__restrict__ dupes) {
  B2N_SKIP_IF(g.skip);
  B2N_GATE(g);
  __shared__ int tkey[HT_N], tcnt[HT_N];
  __shared__ int used;    // lanes merged in this CTA
  __shared__ int filled;  // the table holds entries
  for (int i = threadIdx.x; i < HT_N; i += blockDim.x) { tkey[i] = HT_EMPTY; tcnt[i] = 0; }
  if (threadIdx.x == 0) used = filled = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t Mpad = (M + 31) & ~int64_t(31);
  int merged = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Mpad; i += stride) {
    const bool valid = i < M;
    int key = -2 - lane;  // unique dummy key for tail lanes
    if (valid) {
      T xr, yr, zr;
      fold3(g, i, x, y, z, xr, yr, zr);
      key = point_key(g, xr, yr, zr);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    // any repeated key in the warp = clustered input: then EVERY key of the warp goes through the
    // table (with a few hundred hot keys most of them still appear once per warp)
    const bool hot = __any_sync(0xffffffffu, valid && __popc(peers) > 1);
    if (valid && lane == __ffs(peers) - 1) {
      const int n = __popc(peers);
      bool done = false;
      if (hot) {  // sum it in the CTA's table (3 probes, then straight to L2)
        merged += n - 1;
        int h = ht_slot(key);
        for (int pr = 0; pr < 3 && !done; pr++, h = (h + 1) & (HT_N - 1)) {
          const int old = atomicCAS(&tkey[h], HT_EMPTY, key);
          if (old == HT_EMPTY || old == key) {
            atomicAdd(&tcnt[h], n);
            if (old == HT_EMPTY) filled = 1;
            done = true;
          }
        }
      }
      if (!done) atomicAdd(&hist[key], n);
    }
  }
  if (merged) atomicAdd(&used, merged);  // CTA total first: `dupes` is ONE address for the whole grid
  __syncthreads();
  if (threadIdx.x == 0 && used > 0) atomicAdd(dupes, used);
  if (filled)
    for (int i = threadIdx.x; i < HT_N; i += blockDim.x)
      if (tkey[i] != HT_EMPTY) atomicAdd(&hist[tkey[i]], tcnt[i]);
}

// P1: CTA-local partition of a chunk into buckets of consecutive keys (bucket = key >> shift).
// The chunk's records are staged in shared memory and leave it in bucket order, so that every
// warp store covers a few contiguous runs (full 32-byte sectors, a handful of L2 requests)
// instead of 32 scattered half-sectors.  The first version wrote each record straight to
// tmp[base[bucket] + rank]: 1e8 partial-sector L2 requests, 2.1 ms at C3 (profiles/r01e_sort_*);
// all three sort passes were bound by the rate of such scattered L2 requests, not by DRAM.
//   A  fold, key, rank inside (CTA, bucket) by shared-memory atomic; record -> smem slot
//   B  exclusive scan of the 256 bucket counts; one global atomicAdd per non-empty bucket
//      reserves the CTA's run in that bucket
//   C  dest = off[bucket] + rank: perm[dest] = slot, pbkt[dest] = bucket
//   D  thread t writes output positions t, t + T, ...: tmp[base[b] + j - off[b]] = smem[perm[j]]
constexpr int PT_T = 256;
#ifndef PT_E
#define PT_E 8  // records per thread (float): 2048 per CTA; 4 -> , 6 -> (A/B knob, profiles/r02v)
#endif
template <typename T> struct PartCfg { static constexpr int E = sizeof(T) == 4 ? PT_E : 4; };

// FAST (two-pass sort): there is no histogram yet.  Bucket b owns the fixed region [b * cap,
// (b + 1) * cap) of tmp; a run that does not fit raises *ovf (every later CTA then returns at once
// and the three-pass pipeline takes over), and the fine key histogram is taken here, one RED per
// record, beside the staging -- the separate histogram pass (0.96 ms at C3) repeated this kernel's
// whole per-point instruction stream (loads, fold_rescale, key) just to count.
template <typename T, bool FAST>
__global__ void __launch_bounds__(PT_T) k_partition(SortGeom g, int64_t M, const T *__restrict__ x,
                                                     const T *__restrict__ y,
                                                     const T *__restrict__ z,
                                                     const int32_t *__restrict__ key_start,
                                                     int64_t K, int shift, int nbuckets,
                                                     int32_t *__restrict__ bucket_cur,
                                                     PtRec<T> *__restrict__ tmp, int cap,
                                                     int32_t *__restrict__ hist, int *__restrict__ ovf) {
  B2N_SKIP_IF(g.skip);
  if (FAST) {
    // *ovf changes WHILE this kernel runs: the whole CTA must take one decision (threads that read
    // it at different moments would otherwise part ways in front of a barrier)
    __shared__ int go;
    if (threadIdx.x == 0) go = *reinterpret_cast<volatile int *>(ovf) == 0;
    __syncthreads();
    if (!go) return;
  } else {
    B2N_GATE(g);
  }
  constexpr int E = PartCfg<T>::E, N = PT_T * E;
  __shared__ PtRec<T> srec[N];
  __shared__ unsigned short perm[N];
  __shared__ unsigned char pbkt[N];
  __shared__ int cnt[256], off[256], base[256], wsum[PT_T / 32];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t c0 = (int64_t)blockIdx.x * N;
  const int nloc = (int)min((int64_t)N, M - c0);
  int br[E];  // bucket << 16 | rank inside (CTA, bucket)
  // all of the thread's coordinates are requested BEFORE the first one is used: behind the atomics
  // below the compiler kept each point's loads next to their use -- E dependent DRAM round trips
  // per thread, 64 % of all stall samples on the first FFMA of every fold (profiles/r02v sortprof)
  T cx[E], cy[E], cz[E];
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int sl = e * PT_T + threadIdx.x;
    cx[e] = cy[e] = cz[e] = T(0);
    if (sl < nloc) {
      cx[e] = x[c0 + sl];
      if (g.dim > 1) cy[e] = y[c0 + sl];
      if (g.dim > 2) cz[e] = z[c0 + sl];
    }
  }
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int sl = e * PT_T + threadIdx.x;
    br[e] = -1;
    if (sl < nloc) {
      T xr = fold_rescale(cx[e], g.nf[0]);
      T yr = g.dim > 1 ? fold_rescale(cy[e], g.nf[1]) : T(0);
      T zr = g.dim > 2 ? fold_rescale(cz[e], g.nf[2]) : T(0);
      const int key = point_key(g, xr, yr, zr);
      if (FAST) atomicAdd(&hist[key], 1);
      const int bkt = key >> shift;
      br[e] = (bkt << 16) | atomicAdd(&cnt[bkt], 1);
      srec[sl] = make_rec<T>(xr, yr, zr, c0 + sl);
    }
  }
  __syncthreads();
  {  // B: exclusive scan of cnt[0..255] (one value per thread) + global reservation
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int v = cnt[threadIdx.x];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int pre = 0;
#pragma unroll
    for (int w = 0; w < PT_T / 32; w++) pre += w < wid ? wsum[w] : 0;
    off[threadIdx.x] = pre + inc - v;
    if (threadIdx.x < nbuckets && v > 0) {
      if (FAST) {
        const int at = atomicAdd(&bucket_cur[threadIdx.x], v);
        base[threadIdx.x] = at + v <= cap ? (int)threadIdx.x * cap + at : -1;
        if (at + v > cap) atomicExch(ovf, 1);
      } else {
        const int64_t k0 = (int64_t)threadIdx.x << shift;
        base[threadIdx.x] = key_start[k0 < K ? k0 : K] + atomicAdd(&bucket_cur[threadIdx.x], v);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < E; e++) {
    if (br[e] >= 0) {
      const int b = br[e] >> 16, dest = off[b] + (br[e] & 0xffff);
      perm[dest] = (unsigned short)(e * PT_T + threadIdx.x);
      pbkt[dest] = (unsigned char)b;
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int j = e * PT_T + threadIdx.x;
    if (j < nloc) {
      const int b = pbkt[j];
      if (!FAST || base[b] >= 0) tmp[base[b] + (j - off[b])] = srec[perm[j]];
    }
  }
}

// P2: final placement.  RAW: read the caller's arrays (no P1); else read the partitioned records.
// key_start doubles as the write cursor: pos = atomicAdd(&key_start[key], n) (one L2 round trip
// per distinct key per warp; everything that needs the scan itself -- bin_start, the subproblem
// list -- is derived BEFORE this pass).  Each thread keeps PL_G points in flight so the atomics'
// latency overlaps (all PL_E of them).
constexpr int PL_G = 8;  // points per thread in flight (4: 1.64 ms, 8: 1.44 ms, 16: 2.2 ms at C3)
constexpr int PL_E = 8;  // points per thread; consecutive chunks per CTA keep the L2 window small
// PADDED (two-pass sort): tmp holds bucket b in [b * cap, b * cap + bucket_cnt[b]); a CTA's 2048 slots lie
// inside one bucket (cap is a multiple of 2048), M = end of that bucket's records.
// AGG = false: no search for equal keys inside the warp (MATCH.ANY + the shuffle of the leader's
// base: 41 % of this kernel's stall samples were short-scoreboard waits on the match results).
// Only the two-pass sort with many buckets uses it: there a cluster heavy enough to put equal keys
// into one warp overflows its bucket region first and the three-pass pipeline (which aggregates)
// takes over, so every key of a warp is almost surely distinct and aggregation buys nothing.
template <typename T, bool RAW, bool PADDED = false, bool AGG = true>
__global__ void __launch_bounds__(256) k_place(SortGeom g, int64_t M, const T *__restrict__ x,
                                                const T *__restrict__ y, const T *__restrict__ z,
                                                const PtRec<T> *__restrict__ tmp,
                                                int32_t *__restrict__ key_cursor,
                                                PtRec<T> *__restrict__ out,
                                                const int32_t *__restrict__ dupes, int64_t dupe_limit,
                                                int cap = 0, const int32_t *__restrict__ bucket_cnt = nullptr) {
  B2N_SKIP_IF(g.skip);
  B2N_GATE(g);
  if (dupes && *dupes > dupe_limit) return;  // clustered input: k_place_agg does the work
  const int lane = threadIdx.x & 31;
  int64_t c0 = (int64_t)blockIdx.x * (256 * PL_E);
  if (PADDED) {
    const int b = (int)(c0 / cap);
    M = (int64_t)b * cap + bucket_cnt[b];
  }
  for (int e0 = 0; e0 < PL_E; e0 += PL_G) {
    if (c0 + e0 * 256 + (threadIdx.x - lane) >= M) break;  // whole warp past the end
    PtRec<T> r[PL_G];
    int key[PL_G], base[PL_G], lr[PL_G];  // lr = leader lane | rank among the peers << 8
#pragma unroll
    for (int u = 0; u < PL_G; u++) {
      const int64_t i = c0 + (e0 + u) * 256 + threadIdx.x;
      if (i < M) {
        if (RAW) {
          T xr, yr, zr;
          fold3(g, i, x, y, z, xr, yr, zr);
          r[u] = make_rec<T>(xr, yr, zr, i);
        } else {
          r[u] = tmp[i];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PL_G; u++) {
      const int64_t i = c0 + (e0 + u) * 256 + threadIdx.x;
      key[u] = i < M ? point_key(g, r[u].x, r[u].y, r[u].z) : -1 - lane;  // dummy keys are unique
    }
    if constexpr (!AGG) {
#pragma unroll
      for (int u = 0; u < PL_G; u++) base[u] = key[u] >= 0 ? atomicAdd(&key_cursor[key[u]], 1) : 0;
#pragma unroll
      for (int u = 0; u < PL_G; u++)
        if (key[u] >= 0) out[base[u]] = r[u];
    } else {
#pragma unroll
      for (int u = 0; u < PL_G; u++) {
        const unsigned peers = __match_any_sync(0xffffffffu, key[u]);
        const int leader = __ffs(peers) - 1;
        base[u] = 0;
        if (key[u] >= 0 && lane == leader) base[u] = atomicAdd(&key_cursor[key[u]], __popc(peers));
        lr[u] = leader | (__popc(peers & ((1u << lane) - 1u)) << 8);
      }
#pragma unroll
      for (int u = 0; u < PL_G; u++) {
        const int b = __shfl_sync(0xffffffffu, base[u], lr[u] & 0xff);
        if (key[u] >= 0) out[b + (lr[u] >> 8)] = r[u];
      }
    }
  }
}

// P2, clustered input (dupes > dupe_limit): a CTA takes PA_N consecutive records, sums the
// multiplicity of every key in a shared-memory hash table (each record remembers its table entry
// and its rank inside the CTA), reserves one run per (CTA, key) with a single global atomicAdd,
// and re-reads the records (L2) to place them.  Global atomics per hot key drop from one per warp
// to one per PA_N records.  Keys that do not fit the table fall back to their own global atomic.
constexpr int PA_T = 512, PA_E = 32, PA_N = PA_T * PA_E;  // 16384 records per CTA
constexpr int PA_HT = 4096;
template <typename T, bool RAW>
__global__ void __launch_bounds__(PA_T) k_place_agg(SortGeom g, int64_t M, const T *__restrict__ x,
                                                     const T *__restrict__ y, const T *__restrict__ z,
                                                     const PtRec<T> *__restrict__ tmp,
                                                     int32_t *__restrict__ key_cursor,
                                                     PtRec<T> *__restrict__ out,
                                                     const int32_t *__restrict__ dupes, int64_t dupe_limit) {
  B2N_SKIP_IF(g.skip);
  B2N_GATE(g);
  if (*dupes <= dupe_limit) return;
  extern __shared__ int pa_smem[];
  int *tkey = pa_smem, *tcnt = pa_smem + PA_HT;        // cnt doubles as the reserved base in phase 3
  unsigned *slot = (unsigned *)(pa_smem + 2 * PA_HT);  // per record: entry << 16 | rank (0xffffffff: direct)
  for (int i = threadIdx.x; i < PA_HT; i += PA_T) { tkey[i] = HT_EMPTY; tcnt[i] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t c0 = (int64_t)blockIdx.x * PA_N;
  auto load = [&](int64_t i) {
    if (RAW) {
      T xr, yr, zr;
      fold3(g, i, x, y, z, xr, yr, zr);
      return make_rec<T>(xr, yr, zr, i);
    }
    return tmp[i];
  };
  for (int e = 0; e < PA_E; e++) {
    const int sl = e * PA_T + threadIdx.x;
    const int64_t i = c0 + sl;
    if (i - lane >= M) break;
    int key = -2 - lane;
    if (i < M) {
      PtRec<T> r = load(i);
      key = point_key(g, r.x, r.y, r.z);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1, n = __popc(peers);
    unsigned v = 0xffffffffu;
    if (i < M && lane == leader) {
      int h = (int)(((unsigned)key * 2654435761u) >> 20);  // 12 bits
      for (int pr = 0; pr < 4; pr++, h = (h + 1) & (PA_HT - 1)) {
        const int old = atomicCAS(&tkey[h], HT_EMPTY, key);
        if (old == HT_EMPTY || old == key) {
          const int r0 = atomicAdd(&tcnt[h], n);
          if (r0 + n <= 0xffff) v = ((unsigned)h << 16) | (unsigned)r0; else atomicSub(&tcnt[h], n);
          break;
        }
      }
    }
    v = __shfl_sync(0xffffffffu, v, leader);
    if (i < M) slot[sl] = v == 0xffffffffu ? v : v + (unsigned)__popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PA_HT; i += PA_T)
    if (tkey[i] != HT_EMPTY && tcnt[i] > 0) tcnt[i] = atomicAdd(&key_cursor[tkey[i]], tcnt[i]);
  __syncthreads();
  for (int e = 0; e < PA_E; e++) {
    const int sl = e * PA_T + threadIdx.x;
    const int64_t i = c0 + sl;
    if (i - lane >= M) break;
    PtRec<T> r;
    int key = -2 - lane;
    unsigned v = 0;
    if (i < M) {
      r = load(i);
      v = slot[sl];
      // RAW records are folded afresh: point_key also applies the anchor's periodic shift to them
      if (RAW || v == 0xffffffffu) {
        const int k = point_key(g, r.x, r.y, r.z);
        if (v == 0xffffffffu) key = k;
      }
    }
    // records whose key found no table entry: one global atomic per distinct key per warp
    const unsigned direct = __ballot_sync(0xffffffffu, i < M && v == 0xffffffffu);
    int pos = 0;
    if (direct) {
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      const int leader = __ffs(peers) - 1;
      int b = 0;
      if (key >= 0 && lane == leader) b = atomicAdd(&key_cursor[key], __popc(peers));
      b = __shfl_sync(0xffffffffu, b, leader);
      pos = b + __popc(peers & ((1u << lane) - 1u));
    }
    if (i < M) {
      if (v != 0xffffffffu) pos = tcnt[v >> 16] + (int)(v & 0xffff);
      out[pos] = r;
    }
  }
}

// p[0 .. n) = 0 when *flag == want (the fallback of the two-pass sort starts from clean counters)
__global__ void __launch_bounds__(256) k_zero_if(const int *flag, int want, int32_t *__restrict__ p, int64_t n,
                                                  const int *skip) {
  B2N_SKIP_IF(skip);
  if (*flag != want) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0;
}

// gpu_sort = 0: identity permutation, folded coordinates only
template <typename T>
__global__ void __launch_bounds__(256) k_fold_only(SortGeom g, int64_t M, const T *__restrict__ x,
                                                    const T *__restrict__ y, const T *__restrict__ z,
                                                    PtRec<T> *__restrict__ out) {
  B2N_SKIP_IF(g.skip);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    T xr, yr, zr;
    fold3(g, i, x, y, z, xr, yr, zr);
    out[i] = make_rec<T>(xr, yr, zr, i);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_extract_idx(int64_t M, const PtRec<T> *__restrict__ rec,
                                                      int32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) idx[i] = (int32_t)rec[i].idx;
}

// bin_start[b] = key_start[b * nsub]; subproblems per bin = ceil(count / maxsub)
// (precision_independent.cu:75-81)
__global__ void k_sp_count(int64_t nbins, int nsub, const int32_t *__restrict__ key_start,
                           int maxsub, int32_t *__restrict__ bin_start, int32_t *__restrict__ cnt,
                           const int *skip) {
  B2N_SKIP_IF(skip);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= nbins) {
    const int s = key_start[b * nsub];
    bin_start[b] = s;
    if (b < nbins) cnt[b] = (key_start[(b + 1) * nsub] - s + maxsub - 1) / maxsub;
  }
}
// subproblem -> bin map   (precision_independent.cu:83-91); one thread per subproblem slot finds
// its bin by bisection (a bin with thousands of subproblems -- clustered input -- used to be
// filled by a single thread)
__global__ void k_sp_fill(int64_t nbins, const int32_t *__restrict__ sp_off,
                          int32_t *__restrict__ sp_bin, int64_t cap, const int *skip) {
  B2N_SKIP_IF(skip);
  const int64_t sp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sp >= cap || sp >= sp_off[nbins]) return;
  int64_t lo = 0, hi = nbins;  // largest b with sp_off[b] <= sp
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (sp_off[mid] <= sp) lo = mid; else hi = mid;
  }
  sp_bin[sp] = (int32_t)lo;
}


// ------------------------------------------------------------------------------- setpts cache
// jax-finufft re-sorts the points on every custom call
// (lib/kernels.cc.cu:49-51,64), although forward, JVP and VJP of one step -- and every iteration
// of a solver on a fixed trajectory -- present the same coordinates.  When enabled
// (b2n_set_setpts_cache / B2N_SETPTS_CACHE=1) setpts first folds the coordinate arrays into a
// 64-bit signature (one streaming read, 4*dim*M bytes) and compares it ON THE DEVICE with the
// signature of the point set the plan holds sorted; on a match every sort kernel returns at
// once.  No host round trip, so the call stays stream-ordered and graph-capturable.  The
// signature is an order-dependent sum of avalanche-mixed (coordinate bits, index) words.
__device__ __forceinline__ unsigned long long sig_mix(unsigned long long v) {  // splitmix64 finaliser
  v ^= v >> 30; v *= 0xbf58476d1ce4e5b9ull;
  v ^= v >> 27; v *= 0x94d049bb133111ebull;
  return v ^ (v >> 31);
}
__device__ __forceinline__ unsigned long long sig_bits(float v) { return (unsigned long long)__float_as_uint(v); }
__device__ __forceinline__ unsigned long long sig_bits(double v) { return (unsigned long long)__double_as_longlong(v); }

template <typename T>
__device__ __forceinline__ unsigned long long sig_point(int dim, int64_t i, T x, T y, T z) {
  unsigned long long v = sig_mix(sig_bits(x) + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1));
  if (dim > 1) v = sig_mix(v ^ sig_bits(y));
  if (dim > 2) v = sig_mix(v ^ (sig_bits(z) << 1));
  return v;
}
// VEC: the arrays are 16-byte aligned -- four (float) or two (double) points per load.  The sum is
// commutative, so both variants give the same signature.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) k_pts_signature(int dim, int64_t M, const T *__restrict__ x,
                                                        const T *__restrict__ y, const T *__restrict__ z,
                                                        unsigned long long *__restrict__ acc) {
  __shared__ unsigned long long part[8];
  unsigned long long h = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    constexpr int W = 16 / sizeof(T);
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const int64_t nv = M / W;
    for (int64_t j = tid; j < nv; j += stride) {
      T a[W], b[W], c[W];
      *reinterpret_cast<V *>(a) = __ldcs(reinterpret_cast<const V *>(x) + j);
      if (dim > 1) *reinterpret_cast<V *>(b) = __ldcs(reinterpret_cast<const V *>(y) + j);
      if (dim > 2) *reinterpret_cast<V *>(c) = __ldcs(reinterpret_cast<const V *>(z) + j);
#pragma unroll
      for (int k = 0; k < W; k++) h += sig_point<T>(dim, j * W + k, a[k], dim > 1 ? b[k] : T(0), dim > 2 ? c[k] : T(0));
    }
    for (int64_t i = nv * W + tid; i < M; i += stride)
      h += sig_point<T>(dim, i, x[i], dim > 1 ? y[i] : T(0), dim > 2 ? z[i] : T(0));
  } else {
    for (int64_t i = tid; i < M; i += stride)
      h += sig_point<T>(dim, i, x[i], dim > 1 ? y[i] : T(0), dim > 2 ? z[i] : T(0));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = h;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; w++) t += part[w];
    atomicAdd(acc, t);
  }
}
// sig[0]: accumulator of this setpts (left at zero), sig[1]: signature of the sorted point set,
// sig[2]: the skip flag the sort kernels read
__global__ void k_sig_decide(unsigned long long *sig, unsigned long long salt, int allow) {
  const unsigned long long s = sig[0] ^ salt;
  *reinterpret_cast<int *>(sig + 2) = (allow && s == sig[1]) ? 1 : 0;
  sig[1] = s;
  sig[0] = 0;
}

template <typename U> static int grow(U **p, int64_t *cap, int64_t need, cudaStream_t st) {
  if (need <= *cap) return 0;
  dev_free(*p, st);
  *p = nullptr;
  *cap = 0;
  if (int e = dev_alloc_t(p, (size_t)need, st)) return e;
  *cap = need;
  return 0;
}

template <typename T>
int binsort_points(Plan<T> &p, int64_t M, const T *x, const T *y, const T *z) {
  cudaStream_t st = p.stream;
  PointSet<T> &ps = p.pts;
  ps.M = M;
  if (int e = grow(&ps.rec, &ps.cap_M, std::max<int64_t>(M, 1), st)) return e;
  SortGeom g;
  g.dim = p.dim;
  for (int d = 0; d < 3; d++) {
    g.nf[d] = (int)p.nf[d];
    g.bin[d] = p.bin[d];
    g.nbin[d] = p.nbin[d];
  }
  for (int d = 0; d < 3; d++)
    g.magic[d] = g.bin[d] > 1 ? (unsigned)(((1ull << 32) + g.bin[d] - 1) / g.bin[d]) : 0u;
  g.gate = nullptr;
  g.gate_val = 0;
  g.nsub = ps.nsub = (p.method == 3 && p.dim == 3) ? p.bin[2] : 1;
  g.swr_ns = p.method == 3 ? p.ns : 0;
  // >= 1024 points per CTA: every CTA of the histogram pass sets up (and scans) its hot-key table
  const int nblk = (int)std::min<int64_t>(std::max<int64_t>(cdiv(M, 1024), 1), 148 * 16);
  ps.sorted = p.opts.gpu_sort != 0 || p.method != 1;
  // setpts cache: same M and same sort geometry as the set this plan holds (=> no buffer below is
  // reallocated) and, decided on the device, the same coordinate signature
  g.skip = nullptr;
  const bool was_ok = ps.sig_ok;
  ps.sig_ok = false;
  if (setpts_cache_enabled() && M > 0) {
    if (!ps.sig) {
      if (int e = dev_alloc_t(&ps.sig, 3, st)) return e;
      B2N_CUDA_OK(cudaMemsetAsync(ps.sig, 0, 3 * sizeof(unsigned long long), st));
    }
    unsigned long long salt = 0x243f6a8885a308d3ull ^ (unsigned long long)M;
    const int gw[] = {g.dim, g.nf[0], g.nf[1], g.nf[2], g.bin[0], g.bin[1], g.bin[2], g.nsub, g.swr_ns,
                      p.method, p.maxsub, (int)ps.sorted, (int)sizeof(T)};
    for (int w : gw) salt = (salt ^ (unsigned long long)(unsigned)w) * 0x100000001b3ull;
    const int allow = was_ok && ps.sig_salt == salt;
    ps.sig_salt = salt;
    const bool vec = (((uintptr_t)x | (g.dim > 1 ? (uintptr_t)y : 0) | (g.dim > 2 ? (uintptr_t)z : 0)) & 15) == 0;
    if (vec) k_pts_signature<T, true><<<148 * 8, 256, 0, st>>>(g.dim, M, x, y, z, ps.sig);
    else k_pts_signature<T, false><<<148 * 8, 256, 0, st>>>(g.dim, M, x, y, z, ps.sig);
    k_sig_decide<<<1, 1, 0, st>>>(ps.sig, salt, allow);
    B2N_LAUNCHED(2);
    B2N_LAUNCH_OK();
    g.skip = reinterpret_cast<const int *>(ps.sig + 2);
  }
  // GM kernels on a fine grid that stays in L2 (<= 48 MB): the sort only buys locality the cache
  // already provides, and at these sizes its launches cost more than they save (C1: 1-D, M = N =
  // 1e6, c128: 0.305 ms sorted, 0.260 ms unsorted, reference cuFINUFFT 0.277 ms)
  if (p.method == 1 && (size_t)p.nftot * sizeof(cpx<T>) <= (size_t(48) << 20)) ps.sorted = false;
  if (!ps.sorted) {
    if (M > 0) k_fold_only<T><<<nblk, 256, 0, st>>>(g, M, x, y, z, ps.rec);  B2N_LAUNCHED(1);
    B2N_LAUNCH_OK();
    ps.sp_cap = 0;
    ps.sig_ok = g.skip != nullptr;
    return 0;
  }
  const int64_t nbins = p.nbins;
  const int64_t K = nbins * g.nsub;
  if (K + 1 > 0x7fffffffLL) return B2N_ERR_NDATA_NOTVALID;
  if (K > ps.cap_keys) {
    dev_free(ps.key_cnt, st);
    dev_free(ps.key_start, st);
    ps.key_cnt = ps.key_start = nullptr;
    ps.cap_keys = 0;
    if (int e = dev_alloc_t(&ps.key_cnt, (size_t)K + 8, st)) return e;  // [K]: merged-lane counter of P0
    if (int e = dev_alloc_t(&ps.key_start, (size_t)K + 1, st)) return e;
    ps.cap_keys = K;
  }
  if (nbins > ps.cap_bins) {
    dev_free(ps.bin_start, st);
    dev_free(ps.sp_off, st);
    ps.bin_start = ps.sp_off = nullptr;
    ps.cap_bins = 0;
    if (int e = dev_alloc_t(&ps.bin_start, (size_t)nbins + 1, st)) return e;
    if (int e = dev_alloc_t(&ps.sp_off, (size_t)nbins + 1, st)) return e;
    ps.cap_bins = nbins;
  }
  if (!ps.bucket_cur)
    if (int e = dev_alloc_t(&ps.bucket_cur, 256 + 8, st)) return e;  // [256]: overflow flag of the two-pass sort
  const int64_t sp_cap = std::min<int64_t>(nbins, M) + M / p.maxsub + 1;
  if (int e = grow(&ps.sp_bin, &ps.cap_sp, sp_cap, st)) return e;
  ps.sp_cap = sp_cap;

  int32_t *dupes = ps.key_cnt + K;  // lanes merged by the histogram pass (clustering signal)
  const int64_t dupe_limit = M / 8;
  // bucket geometry: windows of ~4 MB of records, at most 256 buckets; one bucket = no P1
  const int64_t bytes = M * (int64_t)sizeof(PtRec<T>);
  int64_t want = bytes <= (48LL << 20) ? 1 : std::min<int64_t>(256, (bytes + (4LL << 20) - 1) / (4LL << 20));
  int shift = 0;
  while (((K - 1) >> shift) + 1 > want) shift++;
  const int nbuckets = (int)(((K - 1) >> shift) + 1);
  const int nplace = cdiv(std::max<int64_t>(M, 1), 256 * PL_E);
  const size_t pa_smem = (2 * PA_HT + PA_N) * sizeof(int);
  B2N_CUDA_OK(cudaFuncSetAttribute(k_place_agg<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pa_smem));
  B2N_CUDA_OK(cudaFuncSetAttribute(k_place_agg<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pa_smem));
  B2N_CUDA_OK(cudaMemsetAsync(ps.key_cnt, 0, sizeof(int32_t) * (K + 1), st));

  // TWO-PASS SORT (large inputs): P1 partitions into fixed-capacity bucket regions and counts the
  // keys as it goes; the scan and the subproblem list follow; P2 places.  A bucket that outgrows
  // its region (strongly non-uniform input) raises a device flag: the passes of this path then
  // return at once and the three-pass pipeline below runs instead -- all decided on the device.
  static const bool no_fast = getenv("B2N_SORT_THREE_PASS") != nullptr;
  // p.sort_fill < 1: the caller knows that the points fill only that fraction of the slowest axis
  // (type-3 targets sit in the central 1/sigma of the inner grid), i.e. the occupied buckets hold
  // 1/fill times the mean -- without the hint such a set always overflows and pays both pipelines
  const double fill = std::min(1.0, std::max(0.1, p.sort_fill));
  const int64_t cap64 = (((int64_t)((double)M / nbuckets / fill * 1.5) + 8192 + 2047) / 2048) * 2048;
  // What the last attempts taught: a strongly non-uniform set (radial trajectories, clusters) overflows
  // every time, and the abandoned attempt costs ~0.3 ms per 1e8 points.  The flag comes home by an
  // asynchronous copy nobody waits for; whatever value is there when the next setpts is enqueued
  // decides (both pipelines are correct for any input, so a stale value only costs time).  After
  // 15 skipped calls the attempt is made again.
  bool learned_skip = false;
  if (ps.ovf_host) {
    if (ps.fast_hold > 0) {
      ps.fast_hold--;
      learned_skip = true;
    } else if (*reinterpret_cast<volatile int *>(ps.ovf_host) != 0) {
      *reinterpret_cast<volatile int *>(ps.ovf_host) = 0;
      ps.fast_hold = 15;
      learned_skip = true;
    }
  }
  const bool fast = !no_fast && !learned_skip && nbuckets > 1 && cap64 * nbuckets < 0x7fffffffLL;
  const int cap = (int)cap64;
  int *ovf = ps.bucket_cur + 256;
  SortGeom gslow = g, gfast = g;
  if (fast) {
    if (int e = grow(&ps.tmp, &ps.cap_tmp, cap64 * nbuckets, st)) return e;
    B2N_CUDA_OK(cudaMemsetAsync(ps.bucket_cur, 0, sizeof(int32_t) * (256 + 8), st));
    gfast.gate = ovf; gfast.gate_val = 0;
    gslow.gate = ovf; gslow.gate_val = 1;
    k_partition<T, true><<<cdiv(M, PT_T * PartCfg<T>::E), PT_T, 0, st>>>(gfast, M, x, y, z, nullptr, K, shift, nbuckets,
                                                                       ps.bucket_cur, ps.tmp, cap, ps.key_cnt, ovf);
    // overflow: clean counters for the fallback's histogram and partition
    k_zero_if<<<148 * 4, 256, 0, st>>>(ovf, 1, ps.key_cnt, K + 1, g.skip);
    B2N_LAUNCHED(2);
    B2N_LAUNCH_OK();
  }
  if (M > 0) k_key_hist<T><<<nblk, 256, 0, st>>>(gslow, M, x, y, z, ps.key_cnt, dupes);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  if (int e = exclusive_scan_i32(ps.key_cnt, ps.key_start, K, st, g.skip)) return e;

  // subproblem list, entirely on device (the reference reads the total back and syncs,
  // V/src/cuda/3d/spread3d_wrapper.cu:479-487).  Derived from the scan BEFORE the placement pass
  // turns key_start into its write cursors.  key_cnt is free again after the scan and holds the
  // per-bin subproblem counts.
  {
    const int nb_blk = cdiv(nbins + 1, 256);
    k_sp_count<<<nb_blk, 256, 0, st>>>(nbins, g.nsub, ps.key_start, p.maxsub, ps.bin_start, ps.key_cnt, g.skip);  B2N_LAUNCHED(1);
    if (p.method != 1) {  // the GM kernels walk the sorted records directly
      if (int e = exclusive_scan_i32(ps.key_cnt, ps.sp_off, nbins, st, g.skip)) return e;
      k_sp_fill<<<cdiv(sp_cap, 256), 256, 0, st>>>(nbins, ps.sp_off, ps.sp_bin, sp_cap, g.skip);  B2N_LAUNCHED(1);
    }
    B2N_LAUNCH_OK();
  }

  if (M > 0) {
    if (fast) {  // P2 of the two-pass sort: the padded bucket regions, bucket by bucket
      if (nbuckets >= 64)
        k_place<T, false, true, false><<<(unsigned)(cap64 * nbuckets / (256 * PL_E)), 256, 0, st>>>(
            gfast, M, x, y, z, ps.tmp, ps.key_start, ps.rec, nullptr, 0, cap, ps.bucket_cur);
      else
        k_place<T, false, true, true><<<(unsigned)(cap64 * nbuckets / (256 * PL_E)), 256, 0, st>>>(
            gfast, M, x, y, z, ps.tmp, ps.key_start, ps.rec, nullptr, 0, cap, ps.bucket_cur);
      // ... or, after an overflow, P1 + P2 of the three-pass pipeline (write cursors cleared first)
      if (ps.ovf_host) cudaMemcpyAsync(ps.ovf_host, ovf, sizeof(int), cudaMemcpyDeviceToHost, st);
      k_zero_if<<<1, 256, 0, st>>>(ovf, 1, ps.bucket_cur, 256, g.skip);
      B2N_LAUNCHED(2);
    }
    if (nbuckets > 1) {
      if (!fast) {
        if (int e = grow(&ps.tmp, &ps.cap_tmp, M, st)) return e;
        B2N_CUDA_OK(cudaMemsetAsync(ps.bucket_cur, 0, sizeof(int32_t) * 256, st));
      }
      k_partition<T, false><<<cdiv(M, PT_T * PartCfg<T>::E), PT_T, 0, st>>>(gslow, M, x, y, z, ps.key_start, K, shift,
                                                                          nbuckets, ps.bucket_cur, ps.tmp, 0, nullptr, nullptr);
      k_place<T, false><<<nplace, 256, 0, st>>>(gslow, M, x, y, z, ps.tmp, ps.key_start, ps.rec, dupes, dupe_limit);
      k_place_agg<T, false><<<cdiv(M, PA_N), PA_T, pa_smem, st>>>(gslow, M, x, y, z, ps.tmp, ps.key_start, ps.rec, dupes, dupe_limit);
      B2N_LAUNCHED(3);
    } else {
      k_place<T, true><<<nplace, 256, 0, st>>>(g, M, x, y, z, nullptr, ps.key_start, ps.rec, dupes, dupe_limit);
      k_place_agg<T, true><<<cdiv(M, PA_N), PA_T, pa_smem, st>>>(g, M, x, y, z, nullptr, ps.key_start, ps.rec, dupes, dupe_limit);
      B2N_LAUNCHED(2);
    }
    B2N_LAUNCH_OK();
  }
  ps.sig_ok = g.skip != nullptr;
  return 0;
}

// idxnupts for b2n_plan_sort_get: extracted from the records on demand
template <typename T> int materialise_idx(Plan<T> &p) {
  PointSet<T> &ps = p.pts;
  if (!ps.rec) return B2N_ERR_PLAN_NOTVALID;
  if (int e = grow(&ps.idx, &ps.cap_idx, std::max<int64_t>(ps.M, 1), p.stream)) return e;
  if (ps.M > 0) k_extract_idx<T><<<cdiv(ps.M, 256), 256, 0, p.stream>>>(ps.M, ps.rec, ps.idx);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

template int binsort_points<float>(Plan<float> &, int64_t, const float *, const float *, const float *);
template int binsort_points<double>(Plan<double> &, int64_t, const double *, const double *, const double *);
template int materialise_idx<float>(Plan<float> &);
template int materialise_idx<double>(Plan<double> &);

}  // namespace b2n

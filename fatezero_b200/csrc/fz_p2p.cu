// fz_p2p.cu — frame-sharded execution over the GPUs of one NVSwitch box: peer-memory exchange kernels.
//
// The frames of ONE clip are split over the ranks (one process per GPU).  Every op of the UNet forward is per-frame except
//   (1) the sparse-causal self-attention reading K / V of other frames        (attention_register.py:162-193),
//   (2) the joint-frame GroupNorm statistics                                  (resnet.py:338,369; unet_3d_condition.py:439),
//   (3) the temporal Conv1d(k=3) over frames (LoRA down / up, full conv)      (resnet.py:72-78, lora.py:46-54),
//   (4) the temporal attention over frames                                    (models/attention.py:327-337).
// All four are served by ONE primitive over a symmetric arena (a cudaMalloc'd slab per rank, mapped into every peer with CUDA IPC, same
// offsets everywhere): the producer PUSHES 2-D segments straight into the consumers' buffers with 16-byte stores over NVLink and raises
// a flag in the consumer's memory (system-scope release) from the last CTA that finished writing to that consumer; the consumer runs a
// one-warp kernel that spins on its local flags (system-scope acquire) and clears them — it is their only reader, so no sequence
// numbers are needed and the same launch sequence can sit in a CUDA graph.  Every exchange site owns its buffers and flags; a site is
// reused one UNet forward later, by which time every rank has consumed it (each forward contains all-to-all GroupNorm exchanges that
// order all ranks).  No NCCL call sits on the data path.
#include "fz_common.cuh"

#include <algorithm>
#include <cstring>

#include "../../include/fatezero_b200.h"

namespace fz {

constexpr int kP2PMaxSegs = 96;  // 4.7 KiB of kernel parameters (CUDA >= 12.1 allows 32 KiB)
constexpr int kP2PMaxDst = 16;
#ifndef FZ_P2P_TIMEOUT_NS
#define FZ_P2P_TIMEOUT_NS 30000000000ull  // a peer that never arrives becomes a trap after 30 s, not a hung box
#endif

struct P2PSeg {
  const uint8_t* src;
  uint8_t* dst;
  long long src_pitch, dst_pitch;
  int rows, row_bytes;  // row_bytes % 16 == 0
  int dst_slot;         // index into flags / counters, -1 = local copy without a flag
};
struct P2PPushParams {
  P2PSeg seg[kP2PMaxSegs];
  int n_segs;
  int n_dst;
  unsigned* flag[kP2PMaxDst];  // flag word in the DESTINATION rank's arena (peer pointer)
  unsigned* counter;           // local arrival counter (zero between launches)
  unsigned* wait_flags;        // local flag words of this site (32, one per source rank), or null
  unsigned wait_mask;          // sources this rank expects data from at this site
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until_raised(unsigned* flag, const char* what, int who) {
  const uint64_t t0 = global_timer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys(flag) == 0u) {
    if ((++spins & 0x3ff) == 0 && global_timer_ns() - t0 > FZ_P2P_TIMEOUT_NS) {
      printf("fz: %s of peer %d never arrived (frame-sharded exchange)\n", what, who);
      __trap();
    }
  }
  *flag = 0u;  // this thread is the flag's only reader
}

// Exchange = push + wait in ONE launch: every CTA copies its share of the segments into the peers (16-byte stores over NVLink); the CTA
// that arrives last raises the flags of all destinations (system-scope release after a system fence) and then waits for this rank's own
// incoming flags, so that the next kernel of the stream (programmatic dependent launch: griddepcontrol.wait = completion of this grid)
// finds the neighbours' data in place.  Raising before waiting makes the exchange deadlock-free.
__global__ void __launch_bounds__(256) p2p_push_kernel(const __grid_constant__ P2PPushParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const P2PSeg& s = p.seg[blockIdx.y];
  const int vec_per_row = s.row_bytes >> 4;
  const long long total = static_cast<long long>(s.rows) * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row);
    const uint4 v = *reinterpret_cast<const uint4*>(s.src + r * s.src_pitch + (static_cast<long long>(c) << 4));
    *reinterpret_cast<uint4*>(s.dst + r * s.dst_pitch + (static_cast<long long>(c) << 4)) = v;
  }
  __threadfence_system();  // this thread's peer stores are performed before the arrival below
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(p.counter, 1u) == gridDim.x * gridDim.y - 1u;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *p.counter = 0;  // ready for the next launch (stream order)
  __threadfence_system();
  if (threadIdx.x < p.n_dst) st_release_sys(p.flag[threadIdx.x], 1u);
  if (p.wait_flags && threadIdx.x < 32 && ((p.wait_mask >> threadIdx.x) & 1u)) spin_until_raised(p.wait_flags + threadIdx.x, "data", threadIdx.x);
}

// One warp: lane i (bit i of mask) spins until its flag is raised, then clears it.
__global__ void p2p_wait_kernel(unsigned* flags, unsigned mask) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x;
  if (lane < 32 && ((mask >> lane) & 1u)) spin_until_raised(flags + lane, "data", lane);
}

// GroupNorm statistics exchange in ONE single-CTA launch: push this rank's per-image (sum, sumsq) [NB * G] into every peer's inbox slot,
// raise the peers' flags, wait for the peers' statistics, add everything up and leave the total of every statistics set in the slot of its
// first local image (the layout fz_groupnorm_apply_f16 consumes; the other slots of the set are zeroed).
// inbox: [world][NB * G] float2 in this rank's arena (slot `me` unused); peer_inbox[r]: rank r's inbox slot for THIS rank; sums in/out.
struct GnXchgParams {
  float2* peer_inbox[32];
  unsigned* peer_flag[32];
  unsigned* flags;       // local flag words (one per source rank)
  const float2* inbox;   // local inbox
  float2* sums;
  int NB, F_loc, G, world, me;
};
__global__ void gn_combine_kernel(const __grid_constant__ GnXchgParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = p.NB * p.G;
  for (int r = 0; r < p.world; ++r) {
    if (r == p.me) continue;
    for (int i = threadIdx.x; i < n; i += blockDim.x) p.peer_inbox[r][i] = p.sums[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < p.world && threadIdx.x != p.me) {
    st_release_sys(p.peer_flag[threadIdx.x], 1u);
    spin_until_raised(p.flags + threadIdx.x, "GroupNorm statistics", threadIdx.x);
  }
  __syncthreads();
  const int sets = p.NB / p.F_loc, G = p.G, F_loc = p.F_loc;
  for (int i = threadIdx.x; i < sets * G; i += blockDim.x) {
    const int b = i / G, g = i - b * G;
    double sa = 0.0, sb = 0.0;
    for (int f = 0; f < F_loc; ++f) {
      const float2 v = p.sums[(b * F_loc + f) * G + g];
      sa += v.x;
      sb += v.y;
    }
    for (int r = 0; r < p.world; ++r) {
      if (r == p.me) continue;
      const float2* in = p.inbox + static_cast<long long>(r) * n;
      for (int f = 0; f < F_loc; ++f) {
        const float2 v = __ldcg(in + (b * F_loc + f) * G + g);
        sa += v.x;
        sb += v.y;
      }
    }
    // (b, g) is touched by this thread only: no ordering with other threads is needed
    p.sums[(b * F_loc) * G + g] = make_float2(static_cast<float>(sa), static_cast<float>(sb));
    for (int f = 1; f < F_loc; ++f) p.sums[(b * F_loc + f) * G + g] = make_float2(0.f, 0.f);
  }
}

}  // namespace fz

using namespace fz;

extern "C" int fz_p2p_alloc(long long nbytes, void** ptr) {
  FZ_CHECK_ARG(ptr && nbytes > 0, "fz_p2p_alloc: bad args");
  FZ_CUDA(cudaMalloc(ptr, static_cast<size_t>(nbytes)));
  FZ_CUDA(cudaMemset(*ptr, 0, static_cast<size_t>(nbytes)));
  FZ_CUDA(cudaDeviceSynchronize());
  return FZ_OK;
}
extern "C" int fz_p2p_free(void* ptr) {
  FZ_CUDA(cudaFree(ptr));
  return FZ_OK;
}
extern "C" int fz_p2p_export(void* ptr, void* handle64) {
  FZ_CHECK_ARG(ptr && handle64, "fz_p2p_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  FZ_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return FZ_OK;
}
extern "C" int fz_p2p_import(const void* handle64, void** ptr) {
  FZ_CHECK_ARG(ptr && handle64, "fz_p2p_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  FZ_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FZ_OK;
}
extern "C" int fz_p2p_unimport(void* ptr) {
  FZ_CUDA(cudaIpcCloseMemHandle(ptr));
  return FZ_OK;
}

extern "C" int fz_p2p_push(const fz_p2p_seg_t* segs, int n_segs, void* const* flags, void* counter, int n_dst, void* wait_flags, unsigned wait_mask,
                           cudaStream_t stream) {
  FZ_CHECK_ARG(segs && n_segs >= 1 && n_segs <= kP2PMaxSegs && n_dst >= 0 && n_dst <= kP2PMaxDst && counter, "fz_p2p_push: %d segments / %d destinations unsupported",
               n_segs, n_dst);
  P2PPushParams p;
  memset(&p, 0, sizeof(p));
  long long max_vec = 1;
  for (int i = 0; i < n_segs; ++i) {
    const fz_p2p_seg_t& s = segs[i];
    FZ_CHECK_ARG(s.src && s.dst && s.rows > 0 && s.row_bytes > 0 && s.row_bytes % 16 == 0 && s.src_pitch % 16 == 0 && s.dst_pitch % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(s.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(s.dst) & 15) == 0,
                 "fz_p2p_push: segment %d is not 16-byte addressable", i);
    p.seg[i].src = static_cast<const uint8_t*>(s.src); p.seg[i].dst = static_cast<uint8_t*>(s.dst);
    p.seg[i].src_pitch = s.src_pitch; p.seg[i].dst_pitch = s.dst_pitch; p.seg[i].rows = s.rows; p.seg[i].row_bytes = s.row_bytes;
    p.seg[i].dst_slot = s.dst_slot;
    max_vec = std::max(max_vec, static_cast<long long>(s.rows) * (s.row_bytes >> 4));
  }
  p.n_segs = n_segs;
  p.n_dst = n_dst;
  // ~8 vectors per thread; large local re-layouts get up to two waves of CTAs, halo-sized segments a handful
  int gx = static_cast<int>(std::min<long long>(296 / std::max(1, n_segs) + 1, (max_vec + 256 * 8 - 1) / (256 * 8)));
  if (gx < 1) gx = 1;
  for (int d = 0; d < n_dst; ++d) {
    FZ_CHECK_ARG(flags[d], "fz_p2p_push: null flag");
    p.flag[d] = static_cast<unsigned*>(flags[d]);
  }
  p.counter = static_cast<unsigned*>(counter);
  p.wait_flags = static_cast<unsigned*>(wait_flags);
  p.wait_mask = wait_flags ? wait_mask : 0u;
  FZ_CUDA(launch_pdl(p2p_push_kernel, dim3(gx, n_segs), dim3(256), 0, stream, p));
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_p2p_wait(void* flags, unsigned mask, cudaStream_t stream) {
  FZ_CHECK_ARG(flags, "fz_p2p_wait: null pointer");
  if (mask == 0) return FZ_OK;
  FZ_CUDA(launch_pdl(p2p_wait_kernel, dim3(1), dim3(32), 0, stream, static_cast<unsigned*>(flags), mask));
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

extern "C" int fz_gn_combine(void* flags, void* const* peer_flags, void* const* peer_inbox, const void* inbox, void* sums, int NB, int F_loc, int G,
                             int world, int me, cudaStream_t stream) {
  FZ_CHECK_ARG(flags && peer_flags && peer_inbox && inbox && sums && F_loc >= 1 && NB % F_loc == 0 && world >= 1 && world <= 32 && me >= 0 && me < world,
               "fz_gn_combine: bad args");
  GnXchgParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < world; ++r) {
    if (r == me) continue;
    FZ_CHECK_ARG(peer_flags[r] && peer_inbox[r], "fz_gn_combine: null peer pointer");
    p.peer_flag[r] = static_cast<unsigned*>(peer_flags[r]);
    p.peer_inbox[r] = static_cast<float2*>(peer_inbox[r]);
  }
  p.flags = static_cast<unsigned*>(flags); p.inbox = static_cast<const float2*>(inbox); p.sums = static_cast<float2*>(sums);
  p.NB = NB; p.F_loc = F_loc; p.G = G; p.world = world; p.me = me;
  const int n = NB * G;
  const int threads = std::min(1024, std::max(64, (n + 31) / 32 * 32));
  FZ_CUDA(launch_pdl(gn_combine_kernel, dim3(1), dim3(threads), 0, stream, p));
  FZ_CUDA(cudaGetLastError());
  return FZ_OK;
}

// Bulk-copy streaming versions of the per-pixel loss kernels (NCHW logits, all classes of a pixel in one thread).
//
// The register kernels in loss_kernels.cu keep C x VEC logits per thread in flight (8 warps per SM, loads and
// the softmax arithmetic of a warp strictly alternate) and reach ~55 % of the HBM roofline.  Here a tile is TP
// consecutive pixels of one image: its C class rows (and the int64 target row) arrive in a shared-memory stage
// through cp.async.bulk (issued by a dedicated IO warp, completion on an mbarrier), every compute thread handles
// its PPT pixels from shared memory, writes the gradients back IN PLACE and the rows leave through cp.async.bulk
// stores.  Up to four stages (~200 KB per SM) are in flight, independent of the register file.
#pragma once
#include "stream_common.cuh"
#include <stdlib.h>

namespace uda {
namespace pxstream {

using namespace stream;

constexpr size_t kSmemBudget = 220 * 1024;   // stages; coefficient / reduction scratch lives above it
constexpr int kDefaultPpt = 2;
constexpr int kPxTile = 512;                 // pixels per tile = compute threads x pixels per thread
template <int PPT, int TP = kPxTile> constexpr int px_compute_threads() { return TP / PPT; }
template <int PPT, int TP = kPxTile> constexpr int px_threads() { return TP / PPT + 32; }   // + the IO warp

// One tensor map per logit / gradient tensor: [B*C][HW] seen as {256 pixels, HW/256, B*C}, box {256, TP/256, C} — the
// whole [C][TP] tile of a stage is ONE cp.async.bulk.tensor instruction.  With one bulk copy per class row the kernels
// were bound by the per-copy cost (~100 clocks per copy and SM: 96 copies per tile of the consistency kernel, 505 us
// for fp32 and bf16 alike), not by HBM.
struct alignas(64) PxMaps { CUtensorMap in[2]; CUtensorMap out[2]; };

struct PxIO {
  const uint8_t* in[2];       // nten tensors [B,C,HW]
  uint8_t* out[2];            // nout in {0, nten}: gradients, written from the same shared-memory rows
  const long long* target;    // [B,HW] or null
  int nten, nout, B, C, esize, stages, tiles_per_img, tiles_per_cta;
  int use_maps;               // tiles travel through the tensor maps (HW % 256 == 0), else one bulk copy per class row
  long long HW, total_tiles;
  uint32_t stage_bytes;
};

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// NT compute threads (threads 0..NT-1) + one IO warp (threads NT..NT+31).  The compute threads only ever wait for
// a tile to land (`full`); the IO warp waits for a tile to be consumed (`done`), issues its stores and refills the
// stage freed LAG tiles earlier, so neither the store drain nor the copy-issue latency sits on the compute path.
template <int PPT, int NT>
struct PixelPipe {
  static constexpr int TP = NT * PPT;     // pixels per tile: 512, or 256 for the two-tensor kernels (four stages)
  static_assert(TP == kPxTile || TP == kPxTile / 2, "tile size");
  const PxIO& io;
  const PxMaps* maps;
  uint8_t* smem;
  uint32_t full_base, done_base;
  long long t_begin;
  int n_my;
  __device__ PixelPipe(const PxIO& io_, uint8_t* smem_, uint64_t* bars, const PxMaps* maps_ = nullptr)
      : io(io_), maps(maps_), smem(smem_) {
    full_base = smem_u32(bars);
    done_base = full_base + 8u * 4;
    t_begin = (long long)blockIdx.x * io.tiles_per_cta;
    long long t_end = t_begin + io.tiles_per_cta;
    if (t_end > io.total_tiles) t_end = io.total_tiles;
    n_my = t_end > t_begin ? (int)(t_end - t_begin) : 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < io.stages; ++s) { mbar_init(full_base + 8u * s, 1); mbar_init(done_base + 8u * s, NT); }
      fence_barrier_init();
    }
    __syncthreads();
  }
  __device__ __forceinline__ bool is_io() const { return threadIdx.x >= NT; }
  __device__ __forceinline__ int image_of(int k) const { return (int)((t_begin + k) / io.tiles_per_img); }
  __device__ __forceinline__ long long pixel0_of(int k) const { return ((t_begin + k) % io.tiles_per_img) * TP; }
  __device__ __forceinline__ int npix_of(int k) const {
    const long long r = io.HW - pixel0_of(k);
    return (int)(r < TP ? r : TP);
  }
  __device__ __forceinline__ uint8_t* stage(int k) const { return smem + (size_t)(k % io.stages) * io.stage_bytes; }
  __device__ __forceinline__ uint8_t* row(int k, int ten, int c) const {
    return stage(k) + (size_t)(ten * io.C + c) * TP * io.esize;
  }
  __device__ __forceinline__ const long long* target_row(int k) const {
    return reinterpret_cast<const long long*>(stage(k) + (size_t)io.nten * io.C * TP * io.esize);
  }
  // ---- compute threads ----
  __device__ __forceinline__ void wait(int k) const {
    mbar_wait(full_base + 8u * (k % io.stages), (uint32_t)((k / io.stages) & 1));
  }
  __device__ __forceinline__ void release(int k) const {   // tile k consumed / overwritten in place with the outputs
    if (io.nout) fence_async_smem();
    mbar_arrive(done_base + 8u * (k % io.stages));
  }
  // ---- IO warp ----
  // a row is only 1-2 KB, so the copies are issued lane-parallel: one thread issuing ~50 bulk copies per tile
  // serialises on the issue latency (measured: a quarter of the HBM roofline)
  __device__ void load(int k, int lane) const {
    const int b = image_of(k), np = npix_of(k);
    const long long p0 = pixel0_of(k);
    const uint32_t bar = full_base + 8u * (k % io.stages);
    const uint32_t rb = (uint32_t)np * io.esize;
    if (io.use_maps) {     // the whole [C][TP] tile of each tensor in one instruction (out-of-range pixels read as zero)
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)(TP * io.esize) * io.nten * io.C + (io.target ? (uint32_t)np * 8u : 0u));
      __syncwarp();
      if (lane < io.nten) tc::tma_load_3d(smem_u32(row(k, lane, 0)), &maps->in[lane], bar, 0, (int)(p0 / 256), b * io.C);
      if (io.target && lane == 31)
        bulk_load(smem_u32(target_row(k)), io.target + (long long)b * io.HW + p0, (uint32_t)np * 8u, bar);
      return;
    }
    if (lane == 0) mbar_expect_tx(bar, rb * io.nten * io.C + (io.target ? (uint32_t)np * 8u : 0u));
    __syncwarp();
    for (int r = lane; r < io.nten * io.C; r += 32) {
      const int i = r >= io.C ? 1 : 0, c = r - i * io.C;
      bulk_load(smem_u32(row(k, i, c)), (i ? io.in[1] : io.in[0]) + (((long long)b * io.C + c) * io.HW + p0) * io.esize,
                rb, bar);
    }
    if (io.target && lane == 31)
      bulk_load(smem_u32(target_row(k)), io.target + (long long)b * io.HW + p0, (uint32_t)np * 8u, bar);
  }
  __device__ void io_loop() const {
    const int lane = threadIdx.x - NT;
    const int S = io.stages, lag = S >= 4 ? 2 : 1;
    for (int k = 0; k < S && k < n_my; ++k) load(k, lane);
    for (int k = 0; k < n_my; ++k) {
      mbar_wait(done_base + 8u * (k % S), (uint32_t)((k / S) & 1));
      if (io.nout) {
        const int b = image_of(k);
        const long long p0 = pixel0_of(k);
        const uint32_t rb = (uint32_t)npix_of(k) * io.esize;
        if (io.use_maps) {
          if (lane < io.nout) tc::tma_store_3d(&maps->out[lane], smem_u32(row(k, lane, 0)), 0, (int)(p0 / 256), b * io.C);
        } else {
          for (int r = lane; r < io.nout * io.C; r += 32) {
            const int i = r >= io.C ? 1 : 0, c = r - i * io.C;
            bulk_store((i ? io.out[1] : io.out[0]) + (((long long)b * io.C + c) * io.HW + p0) * io.esize,
                       smem_u32(row(k, i, c)), rb);
          }
        }
        bulk_commit();   // bulk groups are per thread: every lane tracks the rows it stored
        // the stage of tile k-lag is free once ITS stores have read shared memory: refill it
        if (k >= lag && k - lag + S < n_my) {
          if (lag == 2) bulk_wait_read<2>(); else bulk_wait_read<1>();
          __syncwarp();
          load(k - lag + S, lane);
        }
      } else if (k + S < n_my) {
        load(k + S, lane);
      }
    }
    if (io.nout) bulk_wait_all<0>();
  }
};

// PPT consecutive pixels of one class row <-> floats
template <typename T, int PPT> __device__ __forceinline__ void ld_px(const uint8_t* rowp, int px, float (&v)[PPT]);
template <> __device__ __forceinline__ void ld_px<float, 1>(const uint8_t* r, int px, float (&v)[1]) {
  v[0] = reinterpret_cast<const float*>(r)[px];
}
template <> __device__ __forceinline__ void ld_px<float, 2>(const uint8_t* r, int px, float (&v)[2]) {
  const float2 t = *reinterpret_cast<const float2*>(r + (size_t)px * 4);
  v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld_px<__nv_bfloat16, 1>(const uint8_t* r, int px, float (&v)[1]) {
  v[0] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(r)[px]);
}
template <> __device__ __forceinline__ void ld_px<__nv_bfloat16, 2>(const uint8_t* r, int px, float (&v)[2]) {
  const uint32_t t = *reinterpret_cast<const uint32_t*>(r + (size_t)px * 2);
  v[0] = __uint_as_float(t << 16); v[1] = __uint_as_float(t & 0xffff0000u);
}
template <typename T, int PPT> __device__ __forceinline__ void st_px(uint8_t* rowp, int px, const float (&v)[PPT]);
template <> __device__ __forceinline__ void st_px<float, 1>(uint8_t* r, int px, const float (&v)[1]) {
  reinterpret_cast<float*>(r)[px] = v[0];
}
template <> __device__ __forceinline__ void st_px<float, 2>(uint8_t* r, int px, const float (&v)[2]) {
  *reinterpret_cast<float2*>(r + (size_t)px * 4) = make_float2(v[0], v[1]);
}
template <> __device__ __forceinline__ void st_px<__nv_bfloat16, 1>(uint8_t* r, int px, const float (&v)[1]) {
  reinterpret_cast<__nv_bfloat16*>(r)[px] = __float2bfloat16_rn(v[0]);
}
template <> __device__ __forceinline__ void st_px<__nv_bfloat16, 2>(uint8_t* r, int px, const float (&v)[2]) {
  *reinterpret_cast<uint32_t*>(r + (size_t)px * 2) = pack_bf16x2(v[0], v[1]);
}

// Host: choose pixels-per-thread and stage count for nten tensors of C classes; false when the streaming path
// does not apply (alignment, too many classes for the shared-memory budget, tiny tensors).
inline bool plan_px(PxIO& io, int& ppt, size_t extra_smem, PxMaps* maps = nullptr, int tile_px = kPxTile,
                    int prefer_ppt = kDefaultPpt) {
  if (io.C > 32 || io.HW <= 0) return false;
  if ((io.HW * io.esize) % 16 != 0 || (io.target && (io.HW * 8) % 16 != 0)) return false;
  for (int i = 0; i < io.nten; ++i)
    if ((reinterpret_cast<uintptr_t>(io.in[i]) & 15) || (io.nout && (reinterpret_cast<uintptr_t>(io.out[i]) & 15))) return false;
  if (io.target && (reinterpret_cast<uintptr_t>(io.target) & 15)) return false;
  if ((long long)io.B * io.C * io.HW * io.esize * io.nten < (1 << 20)) return false;
  (void)extra_smem;
  // pixel pairs per thread (8 compute warps) or single pixels (16 compute warps); UDA_B200_LOSS_PPT overrides
  static const int forced = [] { const char* e = getenv("UDA_B200_LOSS_PPT"); return e ? atoi(e) : 0; }();
  ppt = forced == 1 ? 1 : (forced == 2 ? 2 : prefer_ppt);
  const size_t tp = (size_t)tile_px;
  const size_t sb = (size_t)io.nten * io.C * tp * io.esize + (io.target ? tp * 8 : 0);
  const int st = (int)(kSmemBudget / sb);
  if (st < 2) return false;
  io.stage_bytes = (uint32_t)sb;
  io.stages = st > 4 ? 4 : st;
  io.tiles_per_img = (int)((io.HW + tp - 1) / tp);
  io.total_tiles = (long long)io.B * io.tiles_per_img;
  // every row of a partial tile must stay a multiple of 16 bytes
  if (((io.HW % tp) * io.esize) % 16 != 0) return false;
  const long long grid = io.total_tiles < num_sms() ? io.total_tiles : num_sms();
  io.tiles_per_cta = (int)((io.total_tiles + grid - 1) / grid);
  io.use_maps = 0;
  static const bool maps_on = [] { const char* e = getenv("UDA_B200_LOSS_TMA"); return !(e && e[0] == '0'); }();
  if (maps && maps_on && io.HW % 256 == 0 && tp % 256 == 0 && io.C <= 256) {
    const uint64_t dims[3] = {256, (uint64_t)(io.HW / 256), (uint64_t)io.B * io.C};
    const uint64_t str[2] = {256ull * io.esize, (uint64_t)io.HW * io.esize};
    const uint32_t box[3] = {256, (uint32_t)(tp / 256), (uint32_t)io.C};
    bool ok = true;
    for (int i = 0; i < io.nten && ok; ++i) ok = tc::make_tmap_plain(&maps->in[i], io.esize, io.in[i], 3, dims, str, box);
    for (int i = 0; i < io.nout && ok; ++i) ok = tc::make_tmap_plain(&maps->out[i], io.esize, io.out[i], 3, dims, str, box);
    io.use_maps = ok ? 1 : 0;
  }
  return true;
}
inline unsigned px_grid(const PxIO& io) { return (unsigned)((io.total_tiles + io.tiles_per_cta - 1) / io.tiles_per_cta); }

}  // namespace pxstream
}  // namespace uda

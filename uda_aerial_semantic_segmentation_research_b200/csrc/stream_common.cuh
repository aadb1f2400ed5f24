// Bulk-copy (TMA, cp.async.bulk) streaming pipeline for the HBM-bound elementwise / reduction kernels.
//
// Register-held loads cap the bytes a thread can keep in flight (ncu on the first BN kernels: 16 resident
// warps, long-scoreboard stalls, ~3 TB/s).  Here one elected thread streams contiguous tiles of every
// input tensor into a multi-stage shared-memory ring with cp.async.bulk (completion on an mbarrier), all
// threads compute from shared memory, and results leave through double-buffered shared-memory tiles with
// cp.async.bulk stores — bytes in flight are bounded by shared memory (~100 KB per SM), not registers.
#pragma once
#include "tc_common.cuh"

namespace uda {
namespace stream {

using tc::smem_u32;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mbar_arrive;
using tc::mbar_expect_tx;
using tc::fence_barrier_init;

__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kTileBytes = 8192;            // per tensor per stage: 512 vectors of 16 B
constexpr int kTileVecs = kTileBytes / 16;
constexpr int kThreads = 256;               // 2 vectors per thread per tile

// 16-byte shared-memory vector of 8 bf16 <-> 8 floats
__device__ __forceinline__ void lds8(const uint8_t* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
  v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
}
__device__ __forceinline__ void sts8(uint8_t* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                            pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

}  // namespace stream
}  // namespace uda

// tma.cuh — Tensor Memory Accelerator plumbing: host-side CUtensorMap creation (driver entry
// point fetched through the runtime, so the library does not link libcuda) with a small cache,
// and the device-side cp.async.bulk.tensor wrappers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc05.cuh"

namespace gdl {

// bf16 NHWC activation [N,H,W,C] viewed as a 4-D tensor {C, W, H, N}; box {64, bw, bh, 1},
// 128-byte swizzle, out-of-bounds elements read as zero (that is the convolution padding).
// Returns nullptr on failure (error text set).
const CUtensorMap* tmap_nhwc(const void* ptr, int N, int H, int W, int C, int box_w, int box_h);
// bf16 row-major matrix [rows][K] viewed as {K, rows}; box {64, box_rows}, 128-byte swizzle.
const CUtensorMap* tmap_rows(const void* ptr, int64_t rows, int64_t K, int box_rows);

// Strided {C, W, H, N} view (element strides), box {64, box_w, box_h, box_n}.
const CUtensorMap* tmap_view4(const void* ptr, int C, int W, int H, int N, int64_t sW, int64_t sH, int64_t sN,
                              int box_w, int box_h = 1, int box_n = 1);

// 16-channel tensors (space-to-depth stem input, 32-byte rows): box {16, bw, bh, 1} / {16, box_rows},
// 32-byte swizzle.
const CUtensorMap* tmap_nhwc16(const void* ptr, int N, int H, int W, int box_w, int box_h);
const CUtensorMap* tmap_rows16(const void* ptr, int64_t rows, int64_t K, int box_rows);

}  // namespace gdl

namespace tc05 {

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// Multicast form: the box lands at the same shared-memory offset of every CTA in cta_mask and completes the
// transaction count of the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst_smem, const CUtensorMap* m, uint64_t* bar, int c0,
                                                      int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// cta_group::2 forms: the box lands in THIS CTA's shared memory, the transaction bytes complete on the mbarrier
// given by a shared::cluster address, which may live in the other CTA of the pair (the leader's "full" barrier).
__device__ __forceinline__ void tma2_load_4d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst_smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

// smem -> global tile store (bulk async group); out-of-bounds elements are clipped.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace tc05

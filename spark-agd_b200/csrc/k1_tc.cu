// k1_tc.cu -- K1 for bf16-stored dense shards: margins on the CUDA cores, X^T r on tcgen05 (sm_100a).
//
// north_star's split for the bf16 configuration: the row tile is brought in ONCE by TMA tensor copies
// (128B swizzle) and used twice without ever being widened into registers:
//   phase 1 (CUDA cores, fp64): m_i = x_i . w  -- 256 consumer threads stream the 16-row tile out of
//           shared memory, widen bf16 -> fp64 on the fly and FMA into 4 row accumulators per thread;
//           nothing is retained, so the loop runs at streaming speed;
//   scalar  (1 dedicated warp): margins -> loss', loss (k1_device.cuh); r_i = loss'_i is split into three
//           bf16 pieces (hi / mid / lo, 24 mantissa bits) that form the B operand [N=16 x K=16 rows];
//   phase 2 (tcgen05, fp32 in TMEM): D[c] (128 features x 16) += A (X^T chunk, MN-major view of the SAME
//           swizzled tile) * B, one UTCHMMA per 128-feature chunk, issued by a single thread;
//           every kFlush tiles the accumulators are read back (tcgen05.ld) and added into fp64 registers,
//           so fp32 only ever holds sums over kFlush*16 rows.
// Roles: warps 0-15 consumers, warp 16 TMA producer, warp 17 MMA issuer, warp 18 scalar, warps 20-23 own the fp64
// gradient and flush TMEM; setmaxnreg moves registers from the consumers/aux warps to the flush warpgroup.  Shared-memory ring of 16 KB groups (8 blocks of [16 rows][64
// features]); a group is released by the tcgen05.commit that follows the MMAs reading it.
// Accuracy: margins and losses are fp64-exact like the other kernels; the gradient carries the bf16x3
// split (2^-24) and fp32 partial sums, i.e. ~1e-7 relative (tests/test_gpu_parity.py states the bound).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "agd_common.cuh"
#include "k1_device.cuh"

namespace agd {

namespace {

constexpr int kKR = 16;            // rows per tile = K of one MMA
// RPT = rows per consumer thread.  RPT = 2: 512 consumers (+ 256 aux/flush threads, setmaxnreg 64/56/112 out of the 80 at
// launch).  RPT = 4: 256 consumers, each w value fetched from shared memory serves four rows instead of two -- the kernel is
// bound by shared-memory bandwidth (TMA writes + MMA operand reads + x reads + w reads), and w is the largest reader.
constexpr int kRegsConsumer = 64, kRegsAux = 56, kRegsFlush = 112;
constexpr int kRegsFlush2 = 168;   // two-gradient form: 80 + everything released by 512 consumers (16 each) and 128 aux threads (24 each)
constexpr int kFlush = 8;          // tiles between TMEM -> fp64 flushes (128 rows of fp32 accumulation)
constexpr int kBlockBytes = kKR * 128;     // one [16 rows][64 features] swizzled block

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_tile_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct TcLayout {
  uint32_t ring_off, w_off, b2_off, partial_off, bars_off, tmem_off, total;
};
__host__ __device__ inline TcLayout tc_layout(int ring_groups, int group_bytes, int d) {
  TcLayout L;
  L.ring_off = 0;
  L.w_off = (uint32_t)ring_groups * group_bytes;
  L.b2_off = L.w_off + (uint32_t)d * 8;
  L.partial_off = L.b2_off + 2 * 512;
  L.bars_off = L.partial_off + 2 * kKR * 16 * 8;
  L.tmem_off = L.bars_off + (2 * (uint32_t)ring_groups + 8) * 8;
  L.total = L.tmem_off + 16;
  return L;
}

struct TcArgs {
  const double *labels;
  const double *w;
  const double *w2;  // optional second point (loss only): pass fusion on the bf16 path (fp32-margin mapping)
  double *slabs;
  long long rows;
  int d, kind, slab_stride;
  unsigned long long sample_seed, sample_thresh;
  long long row_base;
  int gb;           // 64-feature blocks per ring group (<= 8)
  int ngt;          // groups per tile = d / (64 * gb)
  int ring_groups;  // ring capacity in groups
  int tmem_cols;    // power of two >= max(32, d / 8)
  int one_copy;     // 1: a ring group arrives as ONE 3-D TMA copy [gb blocks][16 rows][64 features] instead of gb 2-D copies
  int diag;         // option k1_diag: 100 = consumers skip the arithmetic, 101 = the MMAs are not issued (timing bisection only)
};

// RPT = 0 selects the row-per-lane consumer mapping: a warp covers all 16 rows of the tile for two adjacent 8-feature chunks,
// so its w reads are broadcasts (one shared-memory wavefront instead of four) and each lane owns one row's partial dot.
__host__ __device__ constexpr int tc_consumers(int rpt) { return rpt == 4 ? 256 : 512; }
// packed fp32 FMA (SASS FFMA2): d.{x,y} = a.{x,y} * b.{x,y} + c.{x,y}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(uint32_t lo, uint32_t hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// F32: phase 1 in fp32 (option tc_margins=f32, the default): a bf16 is the upper half of an fp32, so widening is one ALU op
// and there is no fp64 conversion per element; products are accumulated by packed fp32 FMAs over at most 8 terms per
// accumulator and then added into the fp64 row sums.  Margins carry ~2^-23 relative to sum |x_i w_i| (w rounded to fp32) --
// the same class as the gradient of this kernel (bf16 x 3 split, fp32 TMEM sums).  F32 = false keeps fp64-exact margins.
// DUAL (F32 mapping only): the loss is also evaluated at a second point w2 from the same tile -- one more packed FMA per
// feature pair in phase 1, lanes 16-31 of the scalar warp -- with bits identical to a launch of its own at w2.
// DUAL == 2: the GRADIENT at w2 as well (two-gradient sweep): r at w2 takes columns 3-5 of the same B operand, so the second
// X^T r costs no extra MMA at all -- the tensor core computes 16 columns either way; the flush reads 8 TMEM columns
// instead of 4 and keeps a second fp64 gradient (setmaxnreg gives the flush warpgroup everything the others released).
template <int RPT, bool F32, int DUAL = 0>
__global__ void __launch_bounds__(tc_consumers(RPT) + 256, 1)
k1_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap3, const TcArgs a,
             const long long ntiles) {
  constexpr int kConsumers = tc_consumers(RPT);   // RPT > 0: 64 threads (one 16-byte vector each) per group row
  constexpr int kThreads = kConsumers + 256;   // + warpgroup (producer, MMA issuer, scalar, idle) + flush warpgroup
  constexpr int kCW = kConsumers / 32;         // consumer warps; the aux warps follow
  constexpr bool kRepartition = RPT != 4;      // 768 threads start with 80 registers: move some to the flush warpgroup
  extern __shared__ __align__(1024) unsigned char smem[];
  const int group_bytes = a.gb * kBlockBytes;
  const TcLayout L = tc_layout(a.ring_groups, group_bytes, a.d);
  double *w_s = reinterpret_cast<double *>(smem + L.w_off);
  unsigned char *b2 = smem + L.b2_off;                                       // [2][512 B]
  double *partial = reinterpret_cast<double *>(smem + L.partial_off);         // [2][16 rows][2] (RPT > 0) or [2][16 rows][16 warps]
  const uint32_t bars = smem_u32(smem + L.bars_off);
  const int RG = a.ring_groups;
  // full[g] = bars + 8g ; empty[g] = bars + 8(RG+g) ; then wbar, b2_full[2], b2_empty[2], tile_done, flush_done
  const uint32_t wbar = bars + 16u * RG, b2_full = wbar + 8, b2_empty = b2_full + 16, tile_done = b2_empty + 16,
                 flush_done = tile_done + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.tmem_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int nch = a.d / 128;
  double *slab = a.slabs + (size_t)blockIdx.x * a.slab_stride;

  if (tid == 0) {
    for (int g = 0; g < RG; ++g) {
      mbar_init(bars + 8u * g, 1);
      mbar_init(bars + 8u * (RG + g), 1);
    }
    mbar_init(wbar, 1);
    mbar_init(b2_full, 1); mbar_init(b2_full + 8, 1);
    mbar_init(b2_empty, 1); mbar_init(b2_empty + 8, 1);
    mbar_init(tile_done, 1);
    mbar_init(flush_done, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // B operand rows 3..15 (unused N columns) stay zero for the whole kernel
  for (int i = tid; i < 2 * 512 / 4; i += kThreads) reinterpret_cast<uint32_t *>(b2)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kCW && warp < kCW + 4) {
   if (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsAux));
   if (warp == kCW) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)a.d * 8u);
      tma_bulk_g2s(smem_u32(w_s), a.w, (uint32_t)a.d * 8u, wbar);
      int slot = -1;
      uint32_t epar = 0;   // parity of the PREVIOUS use of a slot's empty barrier
      bool wrapped = false;
      for (long long k = 0; k < my_tiles; ++k) {
        const long long row0 = (blockIdx.x + k * (long long)gridDim.x) * kKR;
        for (int gi = 0; gi < a.ngt; ++gi) {
          if (++slot == RG) { slot = 0; if (wrapped) epar ^= 1u; wrapped = true; }
          if (wrapped) mbar_wait(bars + 8u * (RG + slot), epar);
          const uint32_t full = bars + 8u * slot;
          mbar_expect_tx(full, (uint32_t)group_bytes);  // rows past the shard are zero-filled by TMA
          if (a.one_copy) {
            tma_tile_3d(smem_u32(smem + (size_t)slot * group_bytes), &tmap3, 0, (int)row0, gi * a.gb, full);
          } else {
            for (int b = 0; b < a.gb; ++b)
              tma_tile_2d(smem_u32(smem + (size_t)slot * group_bytes + b * kBlockBytes), &tmap, (gi * a.gb + b) * 64, (int)row0, full);
          }
        }
      }
    }
   } else if (warp == kCW + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instr desc: D=F32, A=B=BF16, A MN-major, B K-major, N=16, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      int slot = -1;
      uint32_t flush_parity = 0;
      for (long long k = 0; k < my_tiles; ++k) {
        const int bb = (int)(k & 1);
        mbar_wait(b2_full + 8u * bb, (uint32_t)((k >> 1) & 1));
        const bool fresh = (k % kFlush) == 0;  // accumulators were just flushed (or never written)
        if (fresh && k > 0) { mbar_wait(flush_done, flush_parity); flush_parity ^= 1u; }
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint64_t bdesc = (uint64_t)((smem_u32(b2 + bb * 512) & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) |
                               ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
        for (int gi = 0; gi < a.ngt; ++gi) {
          if (++slot == RG) slot = 0;
          for (int cc = 0; cc < a.gb / 2; ++cc) {
            const uint32_t a_addr = smem_u32(smem + (size_t)slot * group_bytes + (2 * cc) * kBlockBytes);
            const uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)(kBlockBytes >> 4) << 16) |
                                   ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t taddr = tmem_base + (uint32_t)(gi * (a.gb / 2) + cc) * 16u;
            const uint32_t acc = fresh ? 0u : 1u;
            if (a.diag != 101)
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(taddr),
                         "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                         : "memory");
          }
          umma_commit(bars + 8u * (RG + slot));  // the group may be refilled once these MMAs have read it
        }
        umma_commit(b2_empty + 8u * bb);
        if (((k + 1) % kFlush) == 0 || k + 1 == my_tiles) umma_commit(tile_done);
      }
    }
   } else if (warp == kCW + 2) {
    // ===================== scalar warp =====================
    // lanes 0-15: the tile's rows at w (loss', loss); DUAL: lanes 16-31 the same rows at w2 (loss only)
    double lossacc = 0.0, cntacc = 0.0;
    double ynext = 0.0;
    const int srow = lane & 15;
    const bool second = DUAL && lane >= kKR;
    if (lane < kKR || DUAL) {
      const long long r = (long long)blockIdx.x * kKR + srow;
      if (r < a.rows) ynext = a.labels[r];
    }
    const double *partial2 = partial + 2 * kKR * 2;   // [2][16 rows][2] at w2 (RPT = 2 layout)
    for (long long k = 0; k < my_tiles; ++k) {
      const int bb = (int)(k & 1);
      const long long tile = blockIdx.x + k * (long long)gridDim.x;
      const long long left = a.rows - tile * kKR;
      const int rv = left < kKR ? (int)left : kKR;
      const double ylab = ynext;
      if (lane < kKR || DUAL) {
        const long long r = (tile + gridDim.x) * kKR + srow;
        if (r < a.rows) ynext = a.labels[r];
      }
      named_sync(1 + bb, kConsumers + 32);                     // partial dots of tile k are in shared memory
      double mult = 0.0;
      if (lane < kKR || second) {
        double m;
        if (RPT) {
          const double *pp = second ? partial2 : partial;
          m = pp[(bb * kKR + srow) * 2] + pp[(bb * kKR + srow) * 2 + 1];
        } else {  // one partial per consumer warp, fixed tree
          const double *pp = partial + (bb * kKR + lane) * 16;
          m = (((pp[0] + pp[1]) + (pp[2] + pp[3])) + ((pp[4] + pp[5]) + (pp[6] + pp[7]))) +
              (((pp[8] + pp[9]) + (pp[10] + pp[11])) + ((pp[12] + pp[13]) + (pp[14] + pp[15])));
        }
        double mu, loss;
        loss_eval(a.kind, m, ylab, mu, loss);
        if (srow < rv && row_selected(a.sample_seed, a.sample_thresh, a.row_base + tile * kKR + srow)) {
          mult = mu; lossacc += loss; cntacc += 1.0;
        }
      }
      named_arrive(3 + bb, kConsumers + 32);                   // partial[bb] may be overwritten
      if (k >= 2) mbar_wait(b2_empty + 8u * bb, (uint32_t)(((k >> 1) - 1) & 1));  // MMAs of tile k-2 have read b2[bb]
      if (lane < kKR || (DUAL == 2 && second)) {
        // r_i -> three bf16 pieces; element (n, row) of the K-major B operand (n = 0..2 at w, 3..5 at w2)
        const __nv_bfloat16 hi = __double2bfloat16(mult);
        const double r1 = mult - (double)__bfloat162float(hi);
        const __nv_bfloat16 mid = __double2bfloat16(r1);
        const double r2 = r1 - (double)__bfloat162float(mid);
        const __nv_bfloat16 lo = __double2bfloat16(r2);
        unsigned char *base = b2 + bb * 512 + (srow / 8) * 128 + (srow % 8) * 2 + (second ? 3 * 16 : 0);
        *reinterpret_cast<__nv_bfloat16 *>(base + 0 * 16) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + 1 * 16) = mid;
        *reinterpret_cast<__nv_bfloat16 *>(base + 2 * 16) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(b2_full + 8u * bb);
    }
    // DUAL: skipping the xor-16 step keeps the lane-0 / lane-16 totals bit-identical to a one-point launch at either point
    // (whose lanes >= 16 only ever contribute exact zeros)
    for (int off = DUAL ? 8 : 16; off >= 1; off >>= 1) {
      lossacc += __shfl_xor_sync(0xffffffffu, lossacc, off);
      cntacc += __shfl_xor_sync(0xffffffffu, cntacc, off);
    }
    if (lane == 0) { slab[a.d] = lossacc; slab[a.d + 1] = cntacc; if (!DUAL) { slab[a.d + 2] = 0.0; slab[a.d + 3] = 0.0; } }
    if (DUAL && lane == kKR) {
      slab[a.d + 2] = lossacc; slab[a.d + 3] = cntacc;
      if (DUAL == 2) {   // second block: [gradient at w2 | loss sum | count | 0 | 0]
        double *slab2 = slab + a.d + 4;
        slab2[a.d] = lossacc; slab2[a.d + 1] = cntacc; slab2[a.d + 2] = 0.0; slab2[a.d + 3] = 0.0;
      }
    }
   }
  } else if (warp >= kCW + 4) {
    // ===================== flush warpgroup: owns the fp64 gradient(s), drains TMEM every kFlush tiles ==========
    if (kRepartition) {
      if (DUAL == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsFlush2));
      else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsFlush));
    }
    const int fw = warp - (kCW + 4);   // TMEM lanes 32*fw .. 32*fw+31 (kCW + 4 is a multiple of 4)
    double gacc[32];                // feature (c*128 + 32*fw + lane), c < d/128
    double gacc2[DUAL == 2 ? 32 : 1];   // the same at w2
#pragma unroll
    for (int c = 0; c < 32; ++c) { gacc[c] = 0.0; if (DUAL == 2) gacc2[DUAL == 2 ? c : 0] = 0.0; }
    uint32_t done_parity = 0;
    for (long long k = 0; k < my_tiles; ++k) {
      if (((k + 1) % kFlush) == 0 || k + 1 == my_tiles) {
        mbar_wait(tile_done, done_parity);
        done_parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c < nch) {
            const uint32_t taddr = tmem_base + ((uint32_t)(fw * 32) << 16) + (uint32_t)c * 16u;
            if (DUAL == 2) {
              uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7)
                           : "r"(taddr));
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              gacc[c] += ((double)__uint_as_float(v0) + (double)__uint_as_float(v1)) + (double)__uint_as_float(v2);
              gacc2[DUAL == 2 ? c : 0] += ((double)__uint_as_float(v3) + (double)__uint_as_float(v4)) + (double)__uint_as_float(v5);
            } else {
              uint32_t v0, v1, v2, v3;
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                           : "r"(taddr));
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              gacc[c] += ((double)__uint_as_float(v0) + (double)__uint_as_float(v1)) + (double)__uint_as_float(v2);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0) mbar_arrive(flush_done);
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c < nch) {
        slab[c * 128 + fw * 32 + lane] = gacc[c];
        if (DUAL == 2) slab[a.d + 4 + c * 128 + fw * 32 + lane] = gacc2[DUAL == 2 ? c : 0];
      }
  } else {
    // ===================== consumers: phase 1 in fp64, straight out of the swizzled tile =====================
    if (kRepartition) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsConsumer));
    if constexpr (RPT == 0) {
      const int r = lane & 15, hsel = lane >> 4;
      mbar_wait(wbar, 0);
      int slot = -1;
      uint32_t par = 1;
      const int npairs = a.gb * 4;                       // pairs of adjacent 8-feature chunks per ring group
      // this lane's chunk in task j: c = 2 * (warp + 16 j) + hsel; block c >> 3, 16-byte position (c & 7) ^ (row & 7)
      uint32_t x_off[2];
      int w_off[2];
      bool act[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 2 * (warp + 16 * j) + hsel;
        act[j] = warp + 16 * j < npairs;
        x_off[j] = (uint32_t)((c >> 3) * kBlockBytes + r * 128 + (((c & 7) ^ (r & 7)) << 4));
        w_off[j] = c * 8;
      }
      for (long long k = 0; k < my_tiles; ++k) {
        const int bb = (int)(k & 1);
        double pa[2] = {0.0, 0.0}, pb[2] = {0.0, 0.0};
        const double *wp = w_s;
        for (int gi = 0; gi < a.ngt; ++gi) {
          if (++slot == RG) slot = 0;
          if (slot == 0) par ^= 1u;
          mbar_wait(bars + 8u * slot, par);
          if (a.diag != 100) {
            const unsigned char *gbase = smem + (uint32_t)slot * (uint32_t)group_bytes;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (act[j]) {
                const uint4 xr = *reinterpret_cast<const uint4 *>(gbase + x_off[j]);
                const double2 *wv = reinterpret_cast<const double2 *>(wp + w_off[j]);   // the same address on 16 lanes
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const double2 wq = wv[q];
                  const uint32_t xw = q == 0 ? xr.x : q == 1 ? xr.y : q == 2 ? xr.z : xr.w;
                  pa[j] = fma((double)__uint_as_float(xw << 16), wq.x, pa[j]);
                  pb[j] = fma((double)__uint_as_float(xw & 0xffff0000u), wq.y, pb[j]);
                }
              }
            }
          }
          wp += a.gb * 64;
        }
        double p = (pa[0] + pb[0]) + (pa[1] + pb[1]);
        p += __shfl_xor_sync(0xffffffffu, p, 16);              // the two chunk columns of the warp
        if (k >= 2) named_sync(3 + bb, kConsumers + 32);         // scalar warp is done with partial[bb] of tile k-2
        if (lane < 16) partial[(bb * kKR + r) * 16 + warp] = p;
        named_arrive(1 + bb, kConsumers + 32);
      }
    } else {
    constexpr int kSlots = kKR / (RPT ? RPT : 1);   // thread (rq, vv) handles rows rq, rq + kSlots, ... of every tile
    const int rq = tid >> 6;
    const int vv = tid & 63;        // 16-byte vector within the group row
    mbar_wait(wbar, 0);
    int slot = -1;
    uint32_t par = 1;               // ring slot and its mbarrier phase, kept incrementally
    const int blk = vv >> 3, ch = vv & 7;
    const bool active = vv < a.gb * 8;
    uint32_t x_off[RPT];            // row rq + j * kSlots of a block: 128-byte rows, 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int row = rq + j * kSlots;
      x_off[j] = (uint32_t)(blk * kBlockBytes + row * 128 + ((ch ^ (row & 7)) << 4));
    }
    if constexpr (F32) {
      // one-time re-layout of w, in place through registers: fp64 (as TMA delivered it) -> fp32 PLANES.  Chunk c (8 features)
      // keeps features 0-3 at plane0[c] and 4-7 at plane1[c] (16 bytes each), so both per-group reads of a warp are
      // contiguous LDS.128s.  The fp64 copy is dead afterwards (the planes overwrite its first half).
      const int nchunks = a.d / 8;
      float4 lo4[2], hi4[2], lo4b[2], hi4b[2];   // d <= 4096: at most 2 chunks per consumer thread
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kConsumers;
        if (c < nchunks) {
          const double *src = w_s + (size_t)c * 8;
          lo4[i] = make_float4((float)src[0], (float)src[1], (float)src[2], (float)src[3]);
          hi4[i] = make_float4((float)src[4], (float)src[5], (float)src[6], (float)src[7]);
          if (DUAL) {   // w2 comes straight from global memory (L2): once per CTA
            const double *s2 = a.w2 + (size_t)c * 8;
            lo4b[i] = make_float4((float)s2[0], (float)s2[1], (float)s2[2], (float)s2[3]);
            hi4b[i] = make_float4((float)s2[4], (float)s2[5], (float)s2[6], (float)s2[7]);
          }
        }
      }
      named_sync(5, kConsumers);
      // planes of w in the first half of the staging area, of w2 in the second half (the fp64 copy of w is dead by now)
      float4 *plane0 = reinterpret_cast<float4 *>(w_s), *plane1 = plane0 + nchunks;
      float4 *plane0b = plane1 + nchunks, *plane1b = plane0b + nchunks;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = tid + i * kConsumers;
        if (c < nchunks) {
          plane0[c] = lo4[i]; plane1[c] = hi4[i];
          if (DUAL) { plane0b[c] = lo4b[i]; plane1b[c] = hi4b[i]; }
        }
      }
      named_sync(5, kConsumers);
      double *partial2 = partial + 2 * kKR * 2;
      for (long long k = 0; k < my_tiles; ++k) {
        const int bb = (int)(k & 1);
        double pd[RPT], pd2[RPT];
        unsigned long long acc[RPT], acc2[RPT];     // packed fp32 pair: even / odd features of this thread's chunk
#pragma unroll
        for (int j = 0; j < RPT; ++j) { pd[j] = 0.0; acc[j] = 0ull; pd2[j] = 0.0; acc2[j] = 0ull; }
        const ulonglong2 *wp0 = reinterpret_cast<const ulonglong2 *>(plane0) + vv;
        const ulonglong2 *wp1 = reinterpret_cast<const ulonglong2 *>(plane1) + vv;
        const ulonglong2 *wq0 = reinterpret_cast<const ulonglong2 *>(plane0b) + vv;
        const ulonglong2 *wq1 = reinterpret_cast<const ulonglong2 *>(plane1b) + vv;
        for (int gi = 0; gi < a.ngt; ++gi) {
          if (++slot == RG) slot = 0;
          if (slot == 0) par ^= 1u;
          mbar_wait(bars + 8u * slot, par);
          if (active && a.diag != 100) {
            const unsigned char *gbase = smem + (uint32_t)slot * (uint32_t)group_bytes;
            uint4 xr[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) xr[j] = *reinterpret_cast<const uint4 *>(gbase + x_off[j]);
            const ulonglong2 wa = *wp0, wb = *wp1;   // (w0,w1),(w2,w3) and (w4,w5),(w6,w7) as packed fp32 pairs
            ulonglong2 va = wa, vb = wb;
            if (DUAL) { va = *wq0; vb = *wq1; }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              const unsigned long long x0 = pack2(xr[j].x << 16, xr[j].x & 0xffff0000u), x1 = pack2(xr[j].y << 16, xr[j].y & 0xffff0000u),
                                       x2 = pack2(xr[j].z << 16, xr[j].z & 0xffff0000u), x3 = pack2(xr[j].w << 16, xr[j].w & 0xffff0000u);
              acc[j] = ffma2(x0, wa.x, acc[j]);
              acc[j] = ffma2(x1, wa.y, acc[j]);
              acc[j] = ffma2(x2, wb.x, acc[j]);
              acc[j] = ffma2(x3, wb.y, acc[j]);
              if (DUAL) {
                acc2[j] = ffma2(x0, va.x, acc2[j]);
                acc2[j] = ffma2(x1, va.y, acc2[j]);
                acc2[j] = ffma2(x2, vb.x, acc2[j]);
                acc2[j] = ffma2(x3, vb.y, acc2[j]);
              }
            }
          }
          wp0 += a.gb * 8;
          wp1 += a.gb * 8;
          wq0 += a.gb * 8;
          wq1 += a.gb * 8;
          if ((gi & 1) || gi + 1 == a.ngt) {   // at most 8 products per fp32 accumulator, then exact fp64
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              pd[j] += (double)(__uint_as_float((uint32_t)acc[j]) + __uint_as_float((uint32_t)(acc[j] >> 32)));
              acc[j] = 0ull;
              if (DUAL) {
                pd2[j] += (double)(__uint_as_float((uint32_t)acc2[j]) + __uint_as_float((uint32_t)(acc2[j] >> 32)));
                acc2[j] = 0ull;
              }
            }
          }
        }
        const double tot = warp_rows_reduce<RPT>(pd, lane);
        double tot2 = 0.0;
        if (DUAL) tot2 = warp_rows_reduce<RPT>(pd2, lane);
        if (k >= 2) named_sync(3 + bb, kConsumers + 32);         // scalar warp is done with partial[bb] of tile k-2
        if ((lane & (32 / RPT - 1)) == 0) {
          partial[(bb * kKR + rq + kSlots * (lane / (32 / RPT))) * 2 + (warp & 1)] = tot;
          if (DUAL) partial2[(bb * kKR + rq + kSlots * (lane / (32 / RPT))) * 2 + (warp & 1)] = tot2;
        }
        named_arrive(1 + bb, kConsumers + 32);
      }
    } else {
    // one-time re-layout of w: within each 64-byte chunk c (8 features) swap the four 16-byte pairs j -> j ^ ((c>>1)&3)
    for (int c = tid; c < a.d / 8; c += kConsumers) {
      const int f = (c >> 1) & 3;
      if (f) {
        double2 *blk = reinterpret_cast<double2 *>(w_s + (size_t)c * 8);
        const double2 t0 = blk[0], t1 = blk[1], t2 = blk[2], t3 = blk[3];
        const double2 v[4] = {t0, t1, t2, t3};
        blk[0 ^ f] = v[0]; blk[1 ^ f] = v[1]; blk[2 ^ f] = v[2]; blk[3 ^ f] = v[3];
      }
    }
    named_sync(5, kConsumers);
    const int sw = (lane >> 1) & 3;
    for (long long k = 0; k < my_tiles; ++k) {
      const int bb = (int)(k & 1);
      double pa[RPT], pb[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) { pa[j] = 0.0; pb[j] = 0.0; }
      const double *wp = w_s + (blk * 64 + ch * 8);
      for (int gi = 0; gi < a.ngt; ++gi) {
        if (++slot == RG) slot = 0;
        if (slot == 0) par ^= 1u;
        mbar_wait(bars + 8u * slot, par);
        if (active && a.diag != 100) {
          const unsigned char *gbase = smem + (uint32_t)slot * (uint32_t)group_bytes;
          uint4 xr[RPT];
#pragma unroll
          for (int j = 0; j < RPT; ++j) xr[j] = *reinterpret_cast<const uint4 *>(gbase + x_off[j]);
          // w was re-laid out once per CTA (above): the 16-byte pair j of chunk c sits at position j ^ ((c >> 1) & 3), so
          // the 8 lanes of a quarter-warp touch 8 different bank groups with no per-element shuffling of the x words
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 wq = *reinterpret_cast<const double2 *>(wp + 2 * (q ^ sw));  // features 2q, 2q+1 of this chunk
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              const uint32_t xw = q == 0 ? xr[j].x : q == 1 ? xr[j].y : q == 2 ? xr[j].z : xr[j].w;
              pa[j] = fma((double)__uint_as_float(xw << 16), wq.x, pa[j]);
              pb[j] = fma((double)__uint_as_float(xw & 0xffff0000u), wq.y, pb[j]);
            }
          }
        }
        wp += a.gb * 64;
      }
      double p[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) p[j] = pa[j] + pb[j];
      // afterwards the lanes of eighth/half-warp j hold the warp total of row rq + j * kSlots
      const double tot = warp_rows_reduce<RPT>(p, lane);
      if (k >= 2) named_sync(3 + bb, kConsumers + 32);         // scalar warp is done with partial[bb] of tile k-2
      if ((lane & (32 / RPT - 1)) == 0) partial[(bb * kKR + rq + kSlots * (lane / (32 / RPT))) * 2 + (warp & 1)] = tot;
      named_arrive(1 + bb, kConsumers + 32);
    }
    }  // fp64 margins
    }  // column-slice mapping
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
  }
}

// the opt-in shared-memory size is a per-device property of a function: set it when it changes, not on every launch
// (one static table per call site, i.e. per kernel instantiation)
template <int SITE, typename K>
cudaError_t set_smem_once(K kern, int bytes) {
  static int cur[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && cur[dev] == bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) cur[dev] = bytes;
  return e;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int k1_tc_supported(int32_t d, int elem_bytes) { return elem_bytes == 2 && d >= 128 && d <= 4096 && d % 128 == 0; }

cudaError_t k1_tc_launch(const K1Args &a, int sm_count, int *blocks_out, cudaStream_t st) {
  if (!k1_tc_supported(a.d, 2)) return cudaErrorInvalidValue;
  if (a.rows <= 0) { *blocks_out = 0; return cudaSuccess; }
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) return e != cudaSuccess ? e : cudaErrorUnknown;
    encode = (EncodeTiledFn)fn;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)a.d, (cuuint64_t)a.rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)a.d * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)kKR};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(a.X), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  CUtensorMap tmap3;
  {
    const cuuint64_t gdim3[3] = {64, (cuuint64_t)a.rows, (cuuint64_t)(a.d / 64)};
    const cuuint64_t gstr3[2] = {(cuuint64_t)a.d * 2, 128};
    const int nblk3 = a.d / 64;
    const cuuint32_t box3[3] = {64, (cuuint32_t)kKR, (cuuint32_t)(nblk3 < 8 ? nblk3 : 8)};
    const cuuint32_t estr3[3] = {1, 1, 1};
    if (encode(&tmap3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(a.X), gdim3, gstr3, box3, estr3,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  TcArgs t;
  t.one_copy = a.tune_ctas == 2 ? 0 : 1;   // option ring_ctas=2: one 2-D copy per 64-feature block (slower: 8x the TMA operations)
  t.diag = (a.kind == 100 || a.kind == 101) ? a.kind : 0;
  if (a.w2 && (a.tune_rows != 0 || a.tc_margins_f64)) return cudaErrorInvalidValue;   // two-point form: default mapping only
  t.labels = a.labels; t.w = a.w; t.w2 = a.w2; t.slabs = a.slabs; t.rows = a.rows; t.d = a.d; t.kind = a.kind;
  t.slab_stride = a.slab_stride;
  t.sample_seed = a.sample_seed; t.sample_thresh = a.sample_thresh; t.row_base = a.row_base;
  const int nblk = a.d / 64;
  t.gb = nblk < 8 ? nblk : 8;
  t.ngt = nblk / t.gb;
  const int group_bytes = t.gb * kBlockBytes;
  int ring = a.stages > 0 ? a.stages : 16;
  const uint32_t budget = 227u * 1024u - 2048u;
  while (ring > 2 && tc_layout(ring, group_bytes, a.d).total + 1024 > budget) --ring;
  if (ring < t.ngt + 1) ring = t.ngt + 1;  // at least one tile and a bit
  t.ring_groups = ring;
  int cols = 32;
  while (cols < a.d / 8) cols <<= 1;
  t.tmem_cols = cols;
  const TcLayout L = tc_layout(ring, group_bytes, a.d);
  const long long ntiles = (a.rows + kKR - 1) / kKR;
  long long grid = sm_count;
  if (grid > ntiles) grid = ntiles;
  *blocks_out = (int)grid;
  const int smem_bytes = (int)L.total + 1024;
  cudaError_t e;
  if (a.tune_rows == 1) {  // option ring_rows=1: row-per-lane consumers (broadcast w reads; measured slower)
    e = set_smem_once<1>(k1_tc_kernel<0, false, 0>, smem_bytes);
    if (e != cudaSuccess) return e;
    k1_tc_kernel<0, false, 0><<<(unsigned)grid, 768, smem_bytes, st>>>(tmap, tmap3, t, ntiles);
  } else if (a.tune_rows == 4) {  // option ring_rows=4: 256 consumers with four rows each (measured slower)
    e = set_smem_once<2>(k1_tc_kernel<4, false, 0>, smem_bytes);
    if (e != cudaSuccess) return e;
    k1_tc_kernel<4, false, 0><<<(unsigned)grid, 512, smem_bytes, st>>>(tmap, tmap3, t, ntiles);
  } else if (a.tc_margins_f64) {  // option tc_margins=f64: 512 consumers, two rows per thread, fp64-exact margins
    e = set_smem_once<3>(k1_tc_kernel<2, false, 0>, smem_bytes);
    if (e != cudaSuccess) return e;
    k1_tc_kernel<2, false, 0><<<(unsigned)grid, 768, smem_bytes, st>>>(tmap, tmap3, t, ntiles);
  } else if (a.w2 && a.dual_full) {  // the default mapping + loss AND gradient at a second point (speculative sweep)
    e = set_smem_once<6>(k1_tc_kernel<2, true, 2>, smem_bytes);
    if (e != cudaSuccess) return e;
    k1_tc_kernel<2, true, 2><<<(unsigned)grid, 768, smem_bytes, st>>>(tmap, tmap3, t, ntiles);
  } else if (a.w2) {  // the default mapping + the loss at a second point (pass fusion)
    e = set_smem_once<4>(k1_tc_kernel<2, true, 1>, smem_bytes);
    if (e != cudaSuccess) return e;
    k1_tc_kernel<2, true, 1><<<(unsigned)grid, 768, smem_bytes, st>>>(tmap, tmap3, t, ntiles);
  } else {  // default: the same mapping with fp32 phase-1 arithmetic (packed FFMA2, no fp64 conversion per element)
    e = set_smem_once<5>(k1_tc_kernel<2, true, 0>, smem_bytes);
    if (e != cudaSuccess) return e;
    k1_tc_kernel<2, true, 0><<<(unsigned)grid, 768, smem_bytes, st>>>(tmap, tmap3, t, ntiles);
  }
  return cudaGetLastError();
}

}  // namespace agd

// tcgen05 implicit-GEMM convolution for sm_100a (tf.nn.conv2d 'SAME' stride 1 + bias + leaky ReLU,
// utils/general.py:36-59), im2col-free:
//
//   D[M = 128 output pixels, N = BN output channels] += A[M, K] * B[N, K]^T,   K = kh*kw*Cin
//
// * A is never materialised.  For filter tap (kh, kw) and 64-channel chunk c the A tile is the 4-D TMA
//   box {64 ch, TW, TH, TB} of the NHWC activation tensor at (c, w0+kw-pad, h0+kh-pad, b0): TMA's
//   out-of-bounds zero fill (negative / past-the-end coordinates) implements the 'SAME' zero padding
//   and never leaks pixels across images.  The box lands in shared memory as 128 rows x 128 bytes with
//   the 128-byte swizzle, which is exactly the canonical K-major SWIZZLE_128B UMMA operand layout.
// * B tiles are 2-D TMA boxes {64, BN} of the pre-packed K-major weights [Cout][kh][kw][Cin].
// * fp32 parity on 16-bit tensor cores: activations and weights are stored as two 16-bit planes
//   x = hi + lo; each K block issues hi*hi + hi*lo + lo*hi (3 passes) into the same fp32 TMEM
//   accumulator (dropped lo*lo term ~2^-18 relative for bf16, ~2^-24 for fp16).  PASSES == 1 is the
//   plain 16-bit path (BASELINE config 5).
// * Warp-specialised persistent CTAs: warp 4 = TMA producer, warp 5 = MMA issuer (+ TMEM alloc),
//   warps 0-3 = epilogue (tcgen05.ld -> fp32 register partial sums -> bias -> leaky ReLU -> hi/lo split or fp32 ->
//   global).  Two TMEM accumulator stages let the epilogue of chunk / tile i overlap the main loop of the next one.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "split_fmt.cuh"

namespace h3d {

namespace {

constexpr int BM = 128;           // UMMA M (pixels per tile)
constexpr int BK = 64;            // K elements per stage (= 128 bytes = one swizzle span)
constexpr int UMMA_K = 16;
constexpr int kNumEpilogueWarps = 4;
constexpr int kThreads = 32 * (kNumEpilogueWarps + 2);
constexpr int kSmemBudget = 227 * 1024 - 2048;
constexpr int A_TILE_BYTES = BM * BK * 2;

// PASSES: 1 = one 16-bit pass; 3 = hi/lo 16-bit planes, three passes; 4 = fp16 plane + two e4m3 planes (each half the bytes),
// one fp16 pass + two fp8 passes.  Modes 3 and 4 stage the same number of bytes.
__host__ __device__ constexpr int stage_bytes(int BN, int PASSES) { return (PASSES >= 3 ? 2 : 1) * (A_TILE_BYTES + BN * BK * 2); }
__host__ __device__ constexpr int num_stages(int BN, int PASSES) {
    return kSmemBudget / stage_bytes(BN, PASSES) > 8 ? 8 : kSmemBudget / stage_bytes(BN, PASSES);
}
__host__ __device__ constexpr int tmem_cols(int BN) { return 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512; }

struct TcParams {
    const float* bias;
    uint16_t* y_hi; uint16_t* y_lo; uint8_t* y_l8; uint8_t* y_h8; int Cy_total, cy_off;
    float corr_scale;   // mode 4: 2^-(10+b), un-does the pre-scaling of the operand planes
    float* yf; int Cyf_total, cyf_off;
    int B, H, W, k, pad, cin_chunks;
    int TW, TH, TB, tiles_w, tiles_h, n_tiles, num_tiles;
    int n_valid;    // number of real output channels (Cout); channels [n_valid, Cout_pad) are padding and never stored
    int pool;       // 1: fuse NetworkOps.max_pool (2x2 / 2) into the epilogue; 2: stride-2 'SAME' convolution (store the odd pixels
                    // of the stride-1 result); outputs are [B, H/2, W/2, C] in both modes
    int chunk_kb;   // K blocks accumulated inside the tensor core before the epilogue folds the partial sum into fp32 registers
    int leaky;
    int stack;      // N-stacked passes (generic single-CTA kernel, 3-pass, BN <= 128); 0 = three separate UMMAs per K step
    int exp;        // reserved for timing experiments (tc_set_tuning("tc_exp")); unused by the shipped kernels
    int tma_out;    // conv_c64x2_kernel: un-pooled split-plane output staged in shared memory and written by bulk tensor stores
    int* err_flag;
    // conv_tc2_kernel, layer chains (see "Layer chains" above the kernel); all optional
    int* sched;            // ticket counter of this launch (zeroed before the launch): dynamic tile scheduler; null = static round-robin
    const int* dep_cnt;    // per-image completion counters of the layer that produces this layer's input; null = griddepcontrol.wait
    int dep_target;        // arrivals per image that mean "every tile of the image has been stored"
    int* sig_cnt;          // this layer's per-image completion counters (zeroed before the launch); null = no signalling
    // conv_c1f_kernel (conv1_1 fused into conv1_2): the first layer's fp32 HWIO weights [3,3,3,64], bias [64] and activation
    const float* w1; const float* bias1; int leaky1;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Suspend-time hint of mbarrier.try_wait: a waiting thread sleeps in hardware (it wakes as soon as the phase completes) instead of
// re-issuing the poll + time-out check every ~20 cycles; ncu r02g: 43 % of the instructions a warp-specialised kernel issued were polls.
constexpr uint32_t kTryWaitHintNs = 20000u;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kTryWaitHintNs) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {   // ~2 s
            if (err_flag) { atomicExch_system(err_flag, code); __threadfence_system(); }
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// TMA prefetch of a tensor box into L2 (no shared-memory destination, no barrier): used to run further ahead of the shared-memory ring
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Programmatic dependent launch: every tensor-core kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization, so its
// CTAs may become resident while the previous kernel of the stream is still draining (on SMs whose CTAs have already exited, or on
// idle SMs when the previous grid is small) and run their prologue - barrier init, TMEM allocation, tensor-map prefetch, resident
// weights - concurrently.  pdl_launch_dependents() lets the NEXT kernel do the same with respect to this one; pdl_wait() blocks until
// every prerequisite grid has completed and its memory is visible, and is executed by every thread before it touches activations
// (reads: TMA producer) or output / workspace buffers (writes: epilogue; ping-pong slots make those WAR-dependent on the previous layer).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One elected lane of a converged warp.  The role warps (TMA producer, MMA issuer) run their loops with ALL 32 lanes in warp-uniform
// control flow and wrap only the issuing instructions in `if (elect_one())`: addresses, descriptors and loop state are then
// warp-uniform values that ptxas keeps in uniform registers, and UTCHMMA / UTMALDG / UTCBAR read their operands straight from them.
// With `if (lane == 0)` around the whole loop (round 1) the operands were per-thread values and every tcgen05.mma was wrapped in a
// divergence "waterfall" (ELECT + 5 R2UR.BROADCAST + BRA.U.ANY, ~20 dependent instructions): ~85 issue cycles per UMMA against 32-64
// tensor cycles for the N = 64 / 128 instructions of the 64-channel kernels, which left their tensor pipe half idle (ncu r02f: the
// issuer never waits on a barrier, yet the pipe computes 50 % of the time).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- CTA-pair (cta_group::2) variants.  Shared-window addresses of a CTA in a cluster carry the CTA rank; clearing the
// peer bit (cute::Sm100MmaPeerBitMask) makes a TMA completion / arrive land on the EVEN CTA's barrier at the same offset.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, void* dst, uint64_t* leader_bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, void* dst, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_f8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at this offset in BOTH CTAs (mask 0b11)
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster.  Default semantics (release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive(cta_id): what these arrivals publish is either nothing in memory (a drained TMEM accumulator, a consumed ring
// slot) or this CTA's OWN shared memory, already made visible to the async proxy by fence.proxy.async, which the tensor core reads on
// behalf of the leader's UMMA.  A `.release.cluster` arrive compiles to MEMBAR.ALL.GPU, which waits for every outstanding global store
// of the thread - the previous tile's output on its way to HBM - and was the top stall of the epilogue warps (ncu r02g: stall_membar).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// ... with release at CLUSTER scope: the arrival publishes data this thread stored into the PEER's shared memory (tile-ticket ring)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// 32-bit store into the shared memory of CTA `rank` of the cluster, at the offset of `p` in this CTA
__device__ __forceinline__ void st_shared_cluster_u32(const void* p, uint32_t rank, uint32_t v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.u32 [ra], %2;\n\t}"
        ::"r"(smem_u32(p)), "r"(rank), "r"(v) : "memory");
}
// wait with cluster-scope acquire: orders shared-memory data written by the PEER CTA before its (release.cluster) arrive
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kTryWaitHintNs) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
    if (mbar_try_wait_cl(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cl(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            if (err_flag) { atomicExch_system(err_flag, code); __threadfence_system(); }
            __trap();
        }
    }
}
// Bounded wait until a device-scope counter reaches `target` (acquire): layer-chain dependencies between two resident grids
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void counter_wait(const int* cnt, int target, int* err_flag, int code) {
    if (ld_acquire_gpu(cnt) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(cnt) < target) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000ll) {
            if (err_flag) { atomicExch_system(err_flag, code); __threadfence_system(); }
            __trap();
        }
    }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version = 1, layout = SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Same for the 8-bit planes: rows of 64 bytes (64 e4m3 values), SWIZZLE_64B (layout 4), SBO = 512 B (8 rows x 64 B).
__device__ __forceinline__ uint64_t make_smem_desc64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format f32 (bit 4), a/b format (bits 7, 10:
// 0 = f16, 1 = bf16), K-major A and B (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int N, bool fp16, int M = BM) {
    return (1u << 4) | ((fp16 ? 0u : 1u) << 7) | ((fp16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <bool FP16>
__device__ __forceinline__ uint32_t pack_hi2(float a, float b) {
    if (FP16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}
template <bool FP16>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
    if (FP16) return __half22float2(*reinterpret_cast<__half2*>(&v));
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

// bias + leaky ReLU + store of 32 consecutive output channels [n, n+32) of one pixel (fp32 and / or hi-lo split planes)
// With p.pool the 2x2 max-pool partners of a pixel are lanes (lane ^ 1) and (lane ^ TW) of the same warp (tile rows are
// ordered w-fastest and TW <= 16), so pooling is two warp shuffles per value; the lane with even (w, h) stores.
template <int PASSES, bool FP16>
__device__ __forceinline__ void epilogue_store32(const TcParams& p, const float* a, int64_t pix, int n, bool valid) {
    float f[32];
    const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 bv = __ldg(bp + q);
        if (PASSES == 4) {   // un-do the operand pre-scaling (exact power of two)
            f[4 * q + 0] = fmaf(a[4 * q + 0], p.corr_scale, bv.x);
            f[4 * q + 1] = fmaf(a[4 * q + 1], p.corr_scale, bv.y);
            f[4 * q + 2] = fmaf(a[4 * q + 2], p.corr_scale, bv.z);
            f[4 * q + 3] = fmaf(a[4 * q + 3], p.corr_scale, bv.w);
        } else {
            f[4 * q + 0] = a[4 * q + 0] + bv.x;
            f[4 * q + 1] = a[4 * q + 1] + bv.y;
            f[4 * q + 2] = a[4 * q + 2] + bv.z;
            f[4 * q + 3] = a[4 * q + 3] + bv.w;
        }
    }
    if (p.leaky) {
#pragma unroll
        for (int q = 0; q < 32; ++q) f[q] = fmaxf(f[q], kNegSlope * f[q]);
    }
    if (p.pool == 1) {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            f[q] = fmaxf(f[q], __shfl_xor_sync(0xFFFFFFFFu, f[q], 1));
            f[q] = fmaxf(f[q], __shfl_xor_sync(0xFFFFFFFFu, f[q], p.TW));
        }
    }
    if (!valid) return;
    if (n + 32 > p.n_valid) {   // Cout not a multiple of 32 (score-map heads 2 / 21, lifting 32-channel layers): masked scalar fp32 tail
        const int cnt = p.n_valid - n;   // <= 0: this 32-channel group is padding only
        if (p.yf && cnt > 0) {
            float* dst = p.yf + pix * p.Cyf_total + p.cyf_off + n;
#pragma unroll
            for (int q = 0; q < 32; ++q)
                if (q < cnt) dst[q] = f[q];
        }
        // the split planes carry Cout_pad channels: padding channels are written as exact zeros (they are the next layer's K padding)
#pragma unroll
        for (int q = 0; q < 32; ++q)
            if (q >= cnt) f[q] = 0.f;
    } else if (p.yf) {
        if (((p.Cyf_total | p.cyf_off) & 3) == 0) {
            float4* dst = reinterpret_cast<float4*>(p.yf + pix * p.Cyf_total + p.cyf_off + n);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else {   // row stride not a multiple of 16 bytes (e.g. the 63-wide fc_xyz output): scalar stores
            float* dst = p.yf + pix * p.Cyf_total + p.cyf_off + n;
#pragma unroll
            for (int q = 0; q < 32; ++q) dst[q] = f[q];
        }
    }
    if (p.y_hi) {
        const int64_t off = pix * p.Cy_total + p.cy_off + n;
        uint4* dh = reinterpret_cast<uint4*>(p.y_hi + off);
        uint4* dl = reinterpret_cast<uint4*>(p.y_lo + off);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float x0 = f[8 * g + 2 * q], x1 = f[8 * g + 2 * q + 1];
                if (PASSES == 4) {   // main plane pre-scaled by 2^5, saturating (must match f32_to_f8c)
                    x0 = fminf(fmaxf(x0 * kF8XMainScale, -65504.f), 65504.f);
                    x1 = fminf(fmaxf(x1 * kF8XMainScale, -65504.f), 65504.f);
                }
                hi[q] = pack_hi2<FP16>(x0, x1);
                if (PASSES == 3) {
                    const float2 r = unpack2<FP16>(hi[q]);
                    lo[q] = pack_hi2<FP16>(x0 - r.x, x1 - r.y);
                }
            }
            dh[g] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (PASSES == 3 && p.y_lo) dl[g] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        if (PASSES == 4) {   // e4m3 residual and coarse planes: 32 bytes each
            uint32_t l8[8], h8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t wl = 0, wh = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const F8cPlanes pl = f32_to_f8c(f[4 * q + e]);
                    wl |= (uint32_t)pl.l8 << (8 * e); wh |= (uint32_t)pl.h8 << (8 * e);
                }
                l8[q] = wl; h8[q] = wh;
            }
            uint4* d8l = reinterpret_cast<uint4*>(p.y_l8 + off);
            uint4* d8h = reinterpret_cast<uint4*>(p.y_h8 + off);
            d8l[0] = make_uint4(l8[0], l8[1], l8[2], l8[3]); d8l[1] = make_uint4(l8[4], l8[5], l8[6], l8[7]);
            d8h[0] = make_uint4(h8[0], h8[1], h8[2], h8[3]); d8h[1] = make_uint4(h8[4], h8[5], h8[6], h8[7]);
        }
    }
}

// ------------------------------------------------------------------------------------------ kernel
template <int BN, int PASSES, bool FP16>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const __grid_constant__ CUtensorMap map_x_h8, const __grid_constant__ CUtensorMap map_w_l8, const TcParams p) {
    // mode 4 operand planes: x_hi = fp16(x), x_lo -> l8 (x residual, e4m3), x_h8 (x, e4m3); w_hi = fp16(w), w_lo -> wh8 (w, e4m3),
    // w_l8 (w residual, e4m3), all pre-scaled so that the three passes fp16 x_hi*w_hi, e4m3 l8*wh8, e4m3 x_h8*w_l8 carry the
    // same power-of-two factor and share one accumulator (split_fmt.cuh); the epilogue multiplies by p.corr_scale.
    constexpr int STAGES = num_stages(BN, PASSES);
    constexpr int STAGE_BYTES = stage_bytes(BN, PASSES);
    constexpr int B_TILE_BYTES = BN * BK * 2;
    constexpr int A8_TILE_BYTES = BM * BK, B8_TILE_BYTES = BN * BK;       // e4m3 tiles: 64-byte rows
    constexpr uint32_t IDESC = make_idesc(BN, FP16);
    // N-stacking (3-pass, BN <= 128): hi*hi and hi*lo are ONE UMMA with N = 2 BN over the adjacent [W_hi ; W_lo] tiles of the stage
    // (TMEM columns [0,BN) and [BN,2BN)), lo*hi a second one with N = BN into columns [0,BN): the A operand is read from shared
    // memory twice instead of three times per K step (the single-CTA kernel is shared-memory-read bound); the epilogue adds the halves.
    constexpr bool STACK = (PASSES == 3) && (BN <= 128);
    constexpr uint32_t IDESC_STACK = make_idesc(STACK ? 2 * BN : BN, FP16);
    constexpr int ACC_COLS = STACK ? 2 * BN : BN;                          // TMEM columns per accumulator stage
    constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;
    static_assert(2 * ACC_COLS <= 512, "TMEM: two accumulator stages must fit 512 columns");
    static_assert(STAGES >= 2, "need at least a double-buffered pipeline");
    static_assert(PASSES != 4 || FP16, "fp8-correction mode uses an fp16 main plane");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.k * p.k * p.cin_chunks;

    if (warp == 4 && lane == 0) {
        prefetch_tmap(&map_x_hi); prefetch_tmap(&map_w_hi);
        if (PASSES >= 3) { prefetch_tmap(&map_x_lo); prefetch_tmap(&map_w_lo); }
        if (PASSES == 4) { prefetch_tmap(&map_x_h8); prefetch_tmap(&map_w_l8); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kNumEpilogueWarps); }

        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 4) {
        // ================================ TMA producer (whole warp, one elected lane issues) ================================
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
            const int w0 = tw * p.TW - p.pad, h0 = th * p.TH - p.pad, b0 = tb * p.TB, n0 = nt * BN;
            int kcol = 0;
            for (int kh = 0; kh < p.k; ++kh) {
                for (int kw = 0; kw < p.k; ++kw) {
                    for (int cc = 0; cc < p.cin_chunks; ++cc, kcol += BK) {
                        mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
                        if (elect_one()) {
                            uint8_t* st = smem + stage * STAGE_BYTES;
                            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                            tma_load_4d(&map_x_hi, st, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                            tma_load_2d(&map_w_hi, st + (PASSES >= 3 ? 2 : 1) * A_TILE_BYTES, &full_bar[stage], kcol, n0);
                            if (PASSES == 3) {
                                tma_load_4d(&map_x_lo, st + A_TILE_BYTES, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                                tma_load_2d(&map_w_lo, st + 2 * A_TILE_BYTES + B_TILE_BYTES, &full_bar[stage], kcol, n0);
                            }
                            if (PASSES == 4) {   // stage = [x fp16 16K | x l8 8K | x h8 8K | w fp16 | w h8 | w l8]
                                tma_load_4d(&map_x_lo, st + A_TILE_BYTES, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                                tma_load_4d(&map_x_h8, st + A_TILE_BYTES + A8_TILE_BYTES, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                                tma_load_2d(&map_w_lo, st + 2 * A_TILE_BYTES + B_TILE_BYTES, &full_bar[stage], kcol, n0);
                                tma_load_2d(&map_w_l8, st + 2 * A_TILE_BYTES + B_TILE_BYTES + B8_TILE_BYTES, &full_bar[stage], kcol, n0);
                            }
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ================================ MMA issuer (whole warp, one elected lane issues) ================================
        int stage = 0; uint32_t phase = 0;
        int acc_it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            for (int kb0 = 0; kb0 < kblocks; kb0 += p.chunk_kb, ++acc_it) {
                const int acc = acc_it & 1;
                mbar_wait(&tempty_bar[acc], ((acc_it >> 1) & 1) ^ 1, p.err_flag, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
                const int kb1 = min(kblocks, kb0 + p.chunk_kb);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t a_hi = make_smem_desc(sa);
                        const uint64_t a_lo = make_smem_desc(sa + A_TILE_BYTES);
                        const uint64_t b_hi = make_smem_desc(sa + (PASSES >= 3 ? 2 : 1) * A_TILE_BYTES);
                        const uint64_t b_lo = make_smem_desc(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
                        for (int j = 0; j < BK / UMMA_K; ++j) {
                            const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);   // advance the start address by 32 B per K step
                            if (STACK && p.stack) {
                                tc_mma_f16(d_tmem, a_hi + koff, b_hi + koff, IDESC_STACK, (uint32_t)((kb > kb0) | (j != 0)));
                                tc_mma_f16(d_tmem, a_lo + koff, b_hi + koff, IDESC, 1u);
                            } else {
                                tc_mma_f16(d_tmem, a_hi + koff, b_hi + koff, IDESC, (uint32_t)((kb > kb0) | (j != 0)));
                                if (PASSES == 3) {
                                    tc_mma_f16(d_tmem, a_hi + koff, b_lo + koff, IDESC, 1u);
                                    tc_mma_f16(d_tmem, a_lo + koff, b_hi + koff, IDESC, 1u);
                                }
                            }
                        }
                        if (PASSES == 4) {   // two e4m3 correction passes (K = 32 per MMA: 32 bytes per row), same accumulator
                            const uint64_t a_l8 = make_smem_desc64(sa + A_TILE_BYTES), a_h8 = make_smem_desc64(sa + A_TILE_BYTES + A8_TILE_BYTES);
                            const uint64_t b_h8 = make_smem_desc64(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
                            const uint64_t b_l8 = make_smem_desc64(sa + 2 * A_TILE_BYTES + B_TILE_BYTES + B8_TILE_BYTES);
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t koff = (uint64_t)((j * 32) >> 4);
                                tc_mma_f8(d_tmem, a_l8 + koff, b_h8 + koff, IDESC, 1u);
                                tc_mma_f8(d_tmem, a_h8 + koff, b_l8 + koff, IDESC, 1u);
                            }
                        }
                        tc_commit(&empty_bar[stage]);   // frees the smem stage once the MMAs above have read it
                        if (kb + 1 == kb1) tc_commit(&tfull_bar[acc]);   // partial accumulator complete -> epilogue (same thread as the MMAs)
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue (warps 0-3 <-> TMEM lanes 32w..32w+31) ================================
        const int row = threadIdx.x;                       // 0..127 = tile row = TMEM lane
        const int w_l = row % p.TW, h_l = (row / p.TW) % p.TH, b_l = row / (p.TW * p.TH);
        int acc_it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
            const int w = tw * p.TW + w_l, h = th * p.TH + h_l, b = tb * p.TB + b_l, n0 = nt * BN;
            bool valid = (w < p.W) && (h < p.H) && (b < p.B);
            int64_t pix = ((int64_t)b * p.H + h) * p.W + w;
            if (p.pool) {   // 1: pooled output pixel, the even-(w, h) lane of each 2x2 window stores; 2: stride-2 'SAME' conv on an
                            // even-sized map = the stride-1 result at the odd pixels (TF pads 0 before / 1 after)
                const int par = p.pool == 2 ? 1 : 0;
                valid = valid && ((w & 1) == par) && ((h & 1) == par);
                pix = ((int64_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            }
            // The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of
            // accumulation steps; K is therefore cut into chunks of chunk_kb blocks whose partial sums are folded
            // into fp32 registers here with round-to-nearest adds (measured: ~10x lower error on K = 4608 layers).
            if constexpr (BN <= 128) {
                float racc[BN];
                for (int kb0 = 0; kb0 < kblocks; kb0 += p.chunk_kb, ++acc_it) {
                    const int acc = acc_it & 1;
                    mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, p.err_flag, 4);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * ACC_COLS);
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        uint32_t v[32];
                        tc_ld_32x32b_x32(taddr + c0, v);
                        tc_wait_ld();
                        if (kb0 == 0) {
#pragma unroll
                            for (int q = 0; q < 32; ++q) racc[c0 + q] = __uint_as_float(v[q]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; ++q) racc[c0 + q] += __uint_as_float(v[q]);
                        }
                        if (STACK && p.stack) {   // + the hi*lo products of the stacked half
                            tc_ld_32x32b_x32(taddr + BN + c0, v);
                            tc_wait_ld();
#pragma unroll
                            for (int q = 0; q < 32; ++q) racc[c0 + q] += __uint_as_float(v[q]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // this warp has drained its 32 lanes of the accumulator
                }
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) epilogue_store32<PASSES, FP16>(p, &racc[c0], pix, n0 + c0, valid);
            } else {
                // BN = 256: 256 fp32 partial sums per thread do not fit the register file -> single TMEM accumulation
                const int acc = acc_it & 1;
                mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, p.err_flag, 4);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    float f[32];
                    tc_ld_32x32b_x32(taddr + c0, v);
                    tc_wait_ld();
#pragma unroll
                    for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]);
                    epilogue_store32<PASSES, FP16>(p, f, pix, n0 + c0, valid);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                ++acc_it;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ 64 -> 64 channel 3x3 kernel
// conv1_2 of both networks (64 -> 64 channels at full resolution, 11 % of the FLOPs) is the one layer shape where the generic
// kernel is bound by L2 -> shared-memory traffic instead of the tensor pipe: with N = 64 every 128-pixel A tile is used for only
// 64 output channels, and the tile is fetched again for each of the 9 filter taps together with the tap's weights (432 KB per
// tile, ~48 B/clk/SM = the measured L2 cap of the chip).  This specialisation removes 72 % of that traffic:
//   * all weights of the layer (9 taps x [W_hi ; W_lo] = 144 KB) are loaded into shared memory ONCE per persistent CTA;
//   * the activations are fetched as three column-shifted patches {64 ch, 16, TH + 2 = 10 rows} per tile (one per kw); the
//     three row taps kh of a patch are the same bytes read through UMMA descriptors whose start address is advanced by
//     kh x 2048 B (= one 16-pixel patch row = two 1024-byte swizzle atoms, so the swizzle phase is unchanged);
//   * N-stacking: hi*hi and hi*lo are ONE UMMA with N = 128 (B = [W_hi ; W_lo], TMEM columns 0-63 / 64-127), lo*hi is a
//     second UMMA with N = 64 into columns 0-63: the A operand is read twice instead of three times (the remaining bound is
//     shared-memory read bandwidth).  The epilogue adds the two column halves in fp32 registers.
// Tiles are fixed at 16 x 8 pixels of one image; everything after the accumulator (bias, leaky ReLU, 2x2 max-pool or
// stride-2 sub-sampling, hi/lo split, masked heads) is the shared epilogue_store32.
constexpr int C64_TW = 16, C64_TH = 8, C64_PH = C64_TH + 2;
constexpr int C64_PATCH_BYTES = C64_PH * C64_TW * BK * 2;                 // 20480: one plane of one patch
constexpr int C64_ROW_BYTES = C64_TW * BK * 2;                            // 2048: one patch row
__host__ __device__ constexpr int c64_w_tap_bytes(int PASSES) { return (PASSES == 3 ? 128 : 64) * BK * 2; }
__host__ __device__ constexpr int c64_a_stage_bytes(int PASSES) { return (PASSES == 3 ? 2 : 1) * C64_PATCH_BYTES; }
__host__ __device__ constexpr int c64_a_stages(int PASSES) { return PASSES == 3 ? 2 : 4; }
__host__ __device__ constexpr int c64_smem_bytes(int PASSES) {
    return 9 * c64_w_tap_bytes(PASSES) + c64_a_stages(PASSES) * c64_a_stage_bytes(PASSES) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <int PASSES, bool FP16>
__global__ void __launch_bounds__(kThreads, 1)
conv_c64_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const TcParams p) {
    static_assert(PASSES == 1 || PASSES == 3, "conv_c64_kernel: one 16-bit pass or the hi/lo 3-pass split");
    constexpr int W_TAP_BYTES = c64_w_tap_bytes(PASSES);
    constexpr int W_BYTES = 9 * W_TAP_BYTES;
    constexpr int A_STAGE_BYTES = c64_a_stage_bytes(PASSES);
    constexpr int STAGES = c64_a_stages(PASSES);
    constexpr int ACC_COLS = PASSES == 3 ? 128 : 64;
    constexpr uint32_t IDESC_MAIN = make_idesc(ACC_COLS, FP16);   // A_hi x [W_hi ; W_lo]  (or A x W for one pass)
    constexpr uint32_t IDESC_N64 = make_idesc(64, FP16);          // A_lo x W_hi
    static_assert(c64_smem_bytes(PASSES) <= 227 * 1024, "conv_c64_kernel: shared memory budget");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem;                                  // [9 taps][W_hi 64 rows | W_lo 64 rows] x 128 B, swizzled
    uint8_t* asm_ = smem + W_BYTES;                       // [STAGES][hi patch | lo patch]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(asm_ + STAGES * A_STAGE_BYTES);
    uint64_t* a_empty = a_full + STAGES;
    uint64_t* tfull_bar = a_empty + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_full = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Cout = 64 * n_tiles: CTA c keeps the weights of the 64-channel group c % n_tiles resident and walks the pixel tiles
    // c / n_tiles, + gridDim / n_tiles, ... (p.num_tiles counts pixel tiles; the grid is a multiple of n_tiles)
    const int n0 = (int)(blockIdx.x % p.n_tiles) * 64;
    const int cta0 = (int)(blockIdx.x / p.n_tiles), cta_step = (int)(gridDim.x / p.n_tiles);

    if (warp == 4 && lane == 0) {
        prefetch_tmap(&map_x_hi); prefetch_tmap(&map_w_hi);
        if (PASSES == 3) { prefetch_tmap(&map_x_lo); prefetch_tmap(&map_w_lo); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kNumEpilogueWarps); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (warp != 4) pdl_wait();   // the TMA producer warp waits after it has issued the (static) weight loads

    if (warp == 4) {
        // ================================ TMA producer (whole warp, one elected lane issues) ================================
        if (elect_one()) {
            mbar_expect_tx(w_full, W_BYTES);
            for (int t = 0; t < 9; ++t) {
                tma_load_2d(&map_w_hi, wsm + t * W_TAP_BYTES, w_full, t * BK, n0);
                if (PASSES == 3) tma_load_2d(&map_w_lo, wsm + t * W_TAP_BYTES + 64 * BK * 2, w_full, t * BK, n0);
            }
        }
        __syncwarp();
        pdl_wait();
        int stage = 0; uint32_t phase = 0;
        for (int tile = cta0; tile < p.num_tiles; tile += cta_step) {
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, tb = tile / (p.tiles_w * p.tiles_h);
            const int w0 = tw * C64_TW - 1, h0 = th * C64_TH - 1;
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&a_empty[stage], phase ^ 1, p.err_flag, 1);
                if (elect_one()) {
                    uint8_t* st = asm_ + stage * A_STAGE_BYTES;
                    mbar_expect_tx(&a_full[stage], A_STAGE_BYTES);
                    tma_load_4d(&map_x_hi, st, &a_full[stage], 0, w0 + kw, h0, tb);
                    if (PASSES == 3) tma_load_4d(&map_x_lo, st + C64_PATCH_BYTES, &a_full[stage], 0, w0 + kw, h0, tb);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        // ================================ MMA issuer (whole warp, one elected lane issues) ================================
        mbar_wait(w_full, 0, p.err_flag, 5);
        tc_fence_after();
        const uint32_t wb = smem_u32(wsm);
        int stage = 0; uint32_t phase = 0;
        int acc_it = 0;
        for (int tile = cta0; tile < p.num_tiles; tile += cta_step, ++acc_it) {
            const int acc = acc_it & 1;
            mbar_wait(&tempty_bar[acc], ((acc_it >> 1) & 1) ^ 1, p.err_flag, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&a_full[stage], phase, p.err_flag, 3);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(asm_ + stage * A_STAGE_BYTES);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint64_t a_hi = make_smem_desc(sa + kh * C64_ROW_BYTES);
                        const uint64_t a_lo = make_smem_desc(sa + C64_PATCH_BYTES + kh * C64_ROW_BYTES);
                        const uint64_t b = make_smem_desc(wb + (kh * 3 + kw) * W_TAP_BYTES);
#pragma unroll
                        for (int j = 0; j < BK / UMMA_K; ++j) {
                            const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                            tc_mma_f16(d_tmem, a_hi + koff, b + koff, IDESC_MAIN, (uint32_t)((kw | kh | j) != 0));
                            if (PASSES == 3) tc_mma_f16(d_tmem, a_lo + koff, b + koff, IDESC_N64, 1u);
                        }
                    }
                    tc_commit(&a_empty[stage]);
                    if (kw == 2) tc_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue (warps 0-3 <-> TMEM lanes 32w..32w+31) ================================
        const int row = threadIdx.x;
        const int w_l = row % C64_TW, h_l = row / C64_TW;
        int acc_it = 0;
        for (int tile = cta0; tile < p.num_tiles; tile += cta_step, ++acc_it) {
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
            const int w = tw * C64_TW + w_l, h = th * C64_TH + h_l;
            bool valid = (w < p.W) && (h < p.H);
            int64_t pix = ((int64_t)b * p.H + h) * p.W + w;
            if (p.pool) {
                const int par = p.pool == 2 ? 1 : 0;
                valid = valid && ((w & 1) == par) && ((h & 1) == par);
                pix = ((int64_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            }
            const int acc = acc_it & 1;
            mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, p.err_flag, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * ACC_COLS);
            float racc[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                tc_ld_32x32b_x32(taddr + c0, v);
                tc_wait_ld();
#pragma unroll
                for (int q = 0; q < 32; ++q) racc[c0 + q] = __uint_as_float(v[q]);
                if (PASSES == 3) {   // + the hi*lo products of the stacked half
                    tc_ld_32x32b_x32(taddr + 64 + c0, v);
                    tc_wait_ld();
#pragma unroll
                    for (int q = 0; q < 32; ++q) racc[c0 + q] += __uint_as_float(v[q]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            epilogue_store32<PASSES, FP16>(p, &racc[0], pix, n0, valid);
            epilogue_store32<PASSES, FP16>(p, &racc[32], pix, n0 + 32, valid);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ACC_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ first layer (Cin = 3) on tensor cores
// conv1_1 of both networks: 3 -> 64 channels, K = 27.  There is nothing for TMA to fetch (3-channel fp32 pixels), so the A
// operand is BUILT in shared memory: four producer warps stage the 18 x 10 x 3 input patch of a 16 x 8 pixel tile, then every
// producer thread writes the 27 neighbourhood values of "its" pixel (+ 5 zeros) as hi / lo 16-bit rows in the canonical
// K-major SWIZZLE_128B layout (byte bits [4,7) ^= bits [7,10); rows keep the 128-byte pitch of the other kernels, only the first
// 64 bytes = 32 K values are ever read).  The 64 x 27 weights are converted and stored the same way once per CTA
// ([W_hi ; W_lo] stacked: 128 rows).  Per tile the issuer runs 2 K steps x { A_hi x [W_hi ; W_lo] (N = 128),
// A_lo x W_hi (N = 64) }; eight epilogue warps add the column halves, apply bias / leaky ReLU and store the split planes.
// The kernel is bound by its 8 bytes / output value of HBM writes (the FFMA kernel it replaces ran at a quarter of that).
constexpr int C3T_THREADS = 32 * 13;                 // warps 0-7 epilogue, 8-11 producers, 12 MMA issuer
constexpr int C3T_STAGES = 2;                        // x 2 CTAs per SM: the kernel is latency bound, not capacity bound
constexpr int C3T_A_STAGE_BYTES = 2 * A_TILE_BYTES;  // hi + lo tile, 128 rows x 128 B each
constexpr int C3T_B_BYTES = 128 * BK * 2;            // [W_hi ; W_lo] x 128 B
constexpr int C3T_PW = C64_TW + 2, C3T_PH = C64_TH + 2;
constexpr int C3T_PATCH_FLOATS = C3T_PH * C3T_PW * 3;   // 540
constexpr int C3T_SMEM = C3T_B_BYTES + C3T_STAGES * C3T_A_STAGE_BYTES + 2 * C3T_PATCH_FLOATS * 4 + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// byte offset of 16-byte chunk c (0..7) of row r in a K-major SWIZZLE_128B tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw128_chunk(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
// Explicit shared-state-space accesses.  The kernels align their dynamic shared memory through an integer round trip, after which the
// compiler no longer knows the address space and emits GENERIC loads / stores (LD.E / ST.E with 64-bit address arithmetic and
// long-scoreboard tracking; ncu r02g) for every `*ptr` access; these take the 32-bit shared window address instead (LDS / STS).
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

template <bool FP16>
__global__ void __launch_bounds__(C3T_THREADS, 2)
conv_c3_tc_kernel(const float* __restrict__ x, const float* __restrict__ w, const TcParams p) {
    constexpr uint32_t IDESC_N128 = make_idesc(128, FP16);
    constexpr uint32_t IDESC_N64 = make_idesc(64, FP16);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bsm = smem;                                   // weights
    uint8_t* asm_ = smem + C3T_B_BYTES;                    // [STAGES][A_hi | A_lo]
    float* patch = reinterpret_cast<float*>(asm_ + C3T_STAGES * C3T_A_STAGE_BYTES);   // [2][PH][PW][3]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(patch + 2 * C3T_PATCH_FLOATS);
    uint64_t* a_empty = a_full + C3T_STAGES;
    uint64_t* tfull_bar = a_empty + C3T_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < C3T_STAGES; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp >= 8 && warp < 12) {
        // ================================ producers: build B once, then one A tile per iteration ================================
        const int t = threadIdx.x - 256;                   // 0..127 = tile row = pixel of the tile
        {   // weights: row n = output channel (t < 64: hi plane, t >= 64: lo plane of channel t - 64), k = (kh*3 + kw)*3 + ci
            const int co = t & 63;
            uint32_t pk[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                float v0 = 0.f, v1 = 0.f;
                if (2 * k2 < 27) v0 = __ldg(w + (2 * k2) * 64 + co);
                if (2 * k2 + 1 < 27) v1 = __ldg(w + (2 * k2 + 1) * 64 + co);
                const uint32_t h = pack_hi2<FP16>(v0, v1);
                if (t < 64) pk[k2] = h;
                else { const float2 r = unpack2<FP16>(h); pk[k2] = pack_hi2<FP16>(v0 - r.x, v1 - r.y); }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(bsm + sw128_chunk(t, c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        const int w_l = t % C64_TW, h_l = t / C64_TW;
        int stage = 0; uint32_t phase = 0; int it = 0;
        // haloed input patch of a tile -> registers (5 independent loads per thread), zero outside the image ('SAME' padding);
        // the loads of tile i+1 are issued before tile i is converted, so their latency is hidden behind the build
        constexpr int PRE = (C3T_PATCH_FLOATS + 127) / 128;
        float pre[PRE];
        auto load_patch = [&](int tile) {
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
            const int x0 = tw * C64_TW - 1, y0 = th * C64_TH - 1;
            const float* xb = x + (int64_t)b * p.H * p.W * 3;
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                const int i = t + j * 128;
                const int r = i / (C3T_PW * 3), rem = i - r * (C3T_PW * 3);
                const int gy = y0 + r, gx = x0 + rem / 3;
                float v = 0.f;
                if (i < C3T_PATCH_FLOATS && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(xb + ((int64_t)gy * p.W + x0) * 3 + rem);
                pre[j] = v;
            }
        };
        if ((int)blockIdx.x < p.num_tiles) load_patch(blockIdx.x);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            float* pb = patch + (it & 1) * C3T_PATCH_FLOATS;
#pragma unroll
            for (int j = 0; j < PRE; ++j)
                if (t + j * 128 < C3T_PATCH_FLOATS) pb[t + j * 128] = pre[j];
            named_bar_sync(1, 128);
            if (tile + (int)gridDim.x < p.num_tiles) load_patch(tile + gridDim.x);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                float v0 = 0.f, v1 = 0.f;
                if (2 * k2 < 27) { const int k = 2 * k2; v0 = pb[(h_l + k / 9) * (C3T_PW * 3) + w_l * 3 + (k % 9)]; }
                if (2 * k2 + 1 < 27) { const int k = 2 * k2 + 1; v1 = pb[(h_l + k / 9) * (C3T_PW * 3) + w_l * 3 + (k % 9)]; }
                hi[k2] = pack_hi2<FP16>(v0, v1);
                const float2 r = unpack2<FP16>(hi[k2]);
                lo[k2] = pack_hi2<FP16>(v0 - r.x, v1 - r.y);
            }
            mbar_wait(&a_empty[stage], phase ^ 1, p.err_flag, 1);
            uint8_t* st = asm_ + stage * C3T_A_STAGE_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<uint4*>(st + sw128_chunk(t, c)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                *reinterpret_cast<uint4*>(st + A_TILE_BYTES + sw128_chunk(t, c)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
            fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core's async-proxy reads
            mbar_arrive(&a_full[stage]);
            if (++stage == C3T_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 12) {
        // ================================ MMA issuer (whole warp, one elected lane issues) ================================
        const uint64_t bdesc = make_smem_desc(smem_u32(bsm));
        int stage = 0; uint32_t phase = 0; int acc_it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++acc_it) {
            const int acc = acc_it & 1;
            mbar_wait(&tempty_bar[acc], ((acc_it >> 1) & 1) ^ 1, p.err_flag, 2);
            mbar_wait(&a_full[stage], phase, p.err_flag, 3);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
                const uint32_t sa = smem_u32(asm_ + stage * C3T_A_STAGE_BYTES);
                const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + A_TILE_BYTES);
#pragma unroll
                for (int j = 0; j < 2; ++j) {                // K = 32 (27 taps x channels + 5 zeros)
                    const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                    tc_mma_f16(d_tmem, a_hi + koff, bdesc + koff, IDESC_N128, (uint32_t)(j != 0));
                    tc_mma_f16(d_tmem, a_lo + koff, bdesc + koff, IDESC_N64, 1u);
                }
                tc_commit(&a_empty[stage]);
                tc_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++stage == C3T_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ================================ epilogue: warps 0-7, lane quadrant = warp % 4, channel half = warp / 4 ================================
        const int q = warp & 3, ch = warp >> 2;
        const int row = q * 32 + lane;
        const int w_l = row % C64_TW, h_l = row / C64_TW;
        int acc_it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++acc_it) {
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
            const int wx = tw * C64_TW + w_l, hy = th * C64_TH + h_l;
            const bool valid = (wx < p.W) && (hy < p.H);
            const int64_t pix = ((int64_t)b * p.H + hy) * p.W + wx;
            const int acc = acc_it & 1;
            mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, p.err_flag, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + ch * 32);
            uint32_t v[32], v2[32];
            tc_ld_32x32b_x32(taddr, v);
            tc_ld_32x32b_x32(taddr + 64, v2);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            float racc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) racc[i] = __uint_as_float(v[i]) + __uint_as_float(v2[i]);
            if (p.y_lo) epilogue_store32<3, FP16>(p, racc, pix, ch * 32, valid);
            else epilogue_store32<1, FP16>(p, racc, pix, ch * 32, valid);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ first layer, TMA-store epilogue
// conv_c3_tc_kernel spends most of its time in the load / store unit: every epilogue thread owns one pixel and writes its 64-byte
// channel slices with 16-byte st.global, so each warp-level store touches 32 different 128-byte lines (half a sector each):
// 2048 LSU cycles per 128-pixel tile (ncu r02b: 2.2 TB/s = 33 % of the HBM roofline, tensor pipe 5 %).  This version stages the
// output tile in shared memory in the SWIZZLE_128B layout (conflict-free 16-byte st.shared) and writes it with ONE bulk tensor
// store per plane (cp.async.bulk.tensor, SASS UTMASTG): full 128-byte lines, 8x fewer LSU cycles, partial tiles clipped by TMA.
// To stay at two CTAs per SM the A operand shrinks: K = 32 needs 64 bytes per row, so A_hi lives in bytes [0,64) and A_lo in
// bytes [64,128) of ONE 128-byte-row tile (start address + 64 B selects the plane, exactly like a K step).
constexpr int C3S_A_STAGE_BYTES = A_TILE_BYTES;          // [A_hi | A_lo] interleaved per row
constexpr int C3S_OUT_BYTES = 2 * A_TILE_BYTES;          // hi plane tile + lo plane tile, 128 rows x 128 B each
constexpr int C3S_SMEM = C3T_B_BYTES + C3T_STAGES * C3S_A_STAGE_BYTES + C3S_OUT_BYTES + 2 * C3T_PATCH_FLOATS * 4 + 1024 + 256;

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <bool FP16>
__global__ void __launch_bounds__(C3T_THREADS, 2)
conv_c3_tma_kernel(const float* __restrict__ x, const float* __restrict__ w, const __grid_constant__ CUtensorMap map_y_hi,
                   const __grid_constant__ CUtensorMap map_y_lo, const TcParams p) {
    constexpr uint32_t IDESC_N128 = make_idesc(128, FP16);
    constexpr uint32_t IDESC_N64 = make_idesc(64, FP16);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* bsm = smem;                                   // weights [W_hi ; W_lo], 128 rows x 128 B (first 64 B used)
    uint8_t* asm_ = smem + C3T_B_BYTES;                    // [STAGES] tiles of 128 rows: bytes [0,64) A_hi, [64,128) A_lo
    uint8_t* osm = asm_ + C3T_STAGES * C3S_A_STAGE_BYTES;  // output staging: [hi tile | lo tile]
    float* patch = reinterpret_cast<float*>(osm + C3S_OUT_BYTES);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(patch + 2 * C3T_PATCH_FLOATS);
    uint64_t* a_empty = a_full + C3T_STAGES;
    uint64_t* tfull_bar = a_empty + C3T_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool want_lo = p.y_lo != nullptr;
    if (threadIdx.x == 0) {
        prefetch_tmap(&map_y_hi);
        if (want_lo) prefetch_tmap(&map_y_lo);
        for (int s = 0; s < C3T_STAGES; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp >= 8 && warp < 12) {
        // ================================ producers: build B once, then one A tile per iteration ================================
        const int t = threadIdx.x - 256;
        {
            const int co = t & 63;
            uint32_t pk[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                float v0 = 0.f, v1 = 0.f;
                if (2 * k2 < 27) v0 = __ldg(w + (2 * k2) * 64 + co);
                if (2 * k2 + 1 < 27) v1 = __ldg(w + (2 * k2 + 1) * 64 + co);
                const uint32_t h = pack_hi2<FP16>(v0, v1);
                if (t < 64) pk[k2] = h;
                else { const float2 r = unpack2<FP16>(h); pk[k2] = pack_hi2<FP16>(v0 - r.x, v1 - r.y); }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(bsm + sw128_chunk(t, c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        const int w_l = t % C64_TW, h_l = t / C64_TW;
        int stage = 0; uint32_t phase = 0; int it = 0;
        constexpr int PRE = (C3T_PATCH_FLOATS + 127) / 128;
        float pre[PRE];
        auto load_patch = [&](int tile) {
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
            const int x0 = tw * C64_TW - 1, y0 = th * C64_TH - 1;
            const float* xb = x + (int64_t)b * p.H * p.W * 3;
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                const int i = t + j * 128;
                const int r = i / (C3T_PW * 3), rem = i - r * (C3T_PW * 3);
                const int gy = y0 + r, gx = x0 + rem / 3;
                float v = 0.f;
                if (i < C3T_PATCH_FLOATS && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(xb + ((int64_t)gy * p.W + x0) * 3 + rem);
                pre[j] = v;
            }
        };
        if ((int)blockIdx.x < p.num_tiles) load_patch(blockIdx.x);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t pb = smem_u32(patch + (it & 1) * C3T_PATCH_FLOATS);
#pragma unroll
            for (int j = 0; j < PRE; ++j)
                if (t + j * 128 < C3T_PATCH_FLOATS) sts_f32(pb + 4u * (uint32_t)(t + j * 128), pre[j]);
            named_bar_sync(1, 128);
            if (tile + (int)gridDim.x < p.num_tiles) load_patch(tile + gridDim.x);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                float v0 = 0.f, v1 = 0.f;
                if (2 * k2 < 27) { const int k = 2 * k2; v0 = lds_f32(pb + 4u * (uint32_t)((h_l + k / 9) * (C3T_PW * 3) + w_l * 3 + (k % 9))); }
                if (2 * k2 + 1 < 27) { const int k = 2 * k2 + 1; v1 = lds_f32(pb + 4u * (uint32_t)((h_l + k / 9) * (C3T_PW * 3) + w_l * 3 + (k % 9))); }
                hi[k2] = pack_hi2<FP16>(v0, v1);
                const float2 r = unpack2<FP16>(hi[k2]);
                lo[k2] = pack_hi2<FP16>(v0 - r.x, v1 - r.y);
            }
            mbar_wait(&a_empty[stage], phase ^ 1, p.err_flag, 1);
            const uint32_t st = smem_u32(asm_ + stage * C3S_A_STAGE_BYTES);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                sts_v4(st + sw128_chunk(t, c), hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                sts_v4(st + sw128_chunk(t, 4 + c), lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&a_full[stage]);
            if (++stage == C3T_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 12) {
        // ================================ MMA issuer (whole warp, one elected lane issues) ================================
        const uint64_t bdesc = make_smem_desc(smem_u32(bsm));
        int stage = 0; uint32_t phase = 0; int acc_it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++acc_it) {
            const int acc = acc_it & 1;
            mbar_wait(&tempty_bar[acc], ((acc_it >> 1) & 1) ^ 1, p.err_flag, 2);
            mbar_wait(&a_full[stage], phase, p.err_flag, 3);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
                const uint64_t a_hi = make_smem_desc(smem_u32(asm_ + stage * C3S_A_STAGE_BYTES));
                const uint64_t a_lo = a_hi + (uint64_t)(64 >> 4);          // bytes [64,128) of every row
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                    tc_mma_f16(d_tmem, a_hi + koff, bdesc + koff, IDESC_N128, (uint32_t)(j != 0));
                    tc_mma_f16(d_tmem, a_lo + koff, bdesc + koff, IDESC_N64, 1u);
                }
                tc_commit(&a_empty[stage]);
                tc_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++stage == C3T_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ================================ epilogue: warps 0-7 (lane quadrant q, channel half ch) -> smem staging -> TMA store ================================
        const int q = warp & 3, ch = warp >> 2;
        const int row = q * 32 + lane;
        int acc_it = 0;
        const uint32_t osm_a = smem_u32(osm);
        const float2* bias2 = reinterpret_cast<const float2*>(p.bias + ch * 32);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++acc_it) {
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
            const int acc = acc_it & 1;
            mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, p.err_flag, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + ch * 32);
            // the previous tile's bulk stores must have finished READING the staging buffer before it is overwritten
            if (warp == 0) { if (elect_one()) tma_store_wait_read(); __syncwarp(); }   // the lane that committed the stores
            named_bar_sync(2, 256);
#pragma unroll
            for (int half = 0; half < 2; ++half) {           // 16 channels at a time (72-register budget at two CTAs per SM)
                uint32_t v[16], v2[16];
                tc_ld_32x32b_x16(taddr + 16 * half, v);          // hi*hi + lo*hi
                tc_ld_32x32b_x16(taddr + 64 + 16 * half, v2);    // hi*lo
                tc_wait_ld();
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 bv = __ldg(bias2 + 8 * half + i);
                    float f0 = (__uint_as_float(v[2 * i]) + __uint_as_float(v2[2 * i])) + bv.x;
                    float f1 = (__uint_as_float(v[2 * i + 1]) + __uint_as_float(v2[2 * i + 1])) + bv.y;
                    if (p.leaky) { f0 = fmaxf(f0, kNegSlope * f0); f1 = fmaxf(f1, kNegSlope * f1); }
                    hi[i] = pack_hi2<FP16>(f0, f1);
                    const float2 r = unpack2<FP16>(hi[i]);
                    lo[i] = pack_hi2<FP16>(f0 - r.x, f1 - r.y);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int chunk = 4 * ch + 2 * half + c;
                    sts_v4(osm_a + sw128_chunk(row, chunk), hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    if (want_lo) sts_v4(osm_a + A_TILE_BYTES + sw128_chunk(row, chunk), lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            fence_proxy_async_smem();
            named_bar_sync(2, 256);
            if (warp == 0) {                                  // warp-uniform; lane 0 is always the elected lane of a full warp
                if (elect_one()) {
                    tma_store_4d(&map_y_hi, osm, 0, tw * C64_TW, th * C64_TH, b);
                    if (want_lo) tma_store_4d(&map_y_lo, osm + A_TILE_BYTES, 0, tw * C64_TW, th * C64_TH, b);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (warp == 0) { if (elect_one()) tma_store_wait_all(); __syncwarp(); }   // global writes complete before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ CTA-pair kernel
// Same algorithm on a cluster of two CTAs (two SMs of one TPC) with tcgen05 cta_group::2: one UMMA covers M = 256 pixels
// (CTA r owns pixel tile 2*pair + r and TMEM rows of it) x N = BN channels; each CTA stages its own A tile and HALF of the
// B tile, so per MMA each SM reads half the shared-memory bytes of the single-CTA kernel (which is shared-memory-bandwidth
// bound at M = N = 128: 8 KB of operands per 64-cycle MMA = the full 128 B/clk) and pulls half the bytes through L2.
// Roles per CTA: warps 0-7 epilogue (lane quadrant = warp % 4, column half = warp / 4), warp 8 TMA producer, warp 9 MMA
// issuer (leader CTA only) + TMEM allocation.  Barriers: full[s] lives in the leader and counts the bytes of both CTAs'
// TMA loads; empty[s] / tmem_full[a] are signalled in both CTAs by multicast commits; tmem_empty[a] lives in the leader
// and collects one arrive per epilogue warp of both CTAs.
constexpr int kThreads2 = 32 * 10;
// N-stacked CTA-pair variant (3-pass, BN == 128): per K step  A_hi x [W_hi ; W_lo]  (one UMMA 256 x 256: TMEM columns [0,128) = hi*hi,
// [128,256) = hi*lo) and  A_lo x W_hi  (UMMA 256 x 128 into columns [0,128)).  With cta_group::2 every CTA supplies half of the B
// rows of an instruction, read at the SAME shared-memory offset in both CTAs, so each stage holds two weight regions:
//   Y (128 rows): CTA 0 = W_hi[n0 .. n0+128), CTA 1 = W_lo[n0 .. n0+128)   -> B operand of the N = 256 instruction
//   X ( 64 rows): CTA r = W_hi[n0 + 64 r .. + 64)                          -> B operand of the N = 128 instruction
// Per K step each SM reads 14 KB of operands in 192 tensor cycles (73 B/clk; the single-CTA stacked kernel reads 20 KB = 107 B/clk,
// which together with the TMA writes exceeds the 128 B/clk of shared memory) and pulls 56 KB per K block through L2 instead of 64.
__host__ __device__ constexpr bool stack2(int BN, int PASSES) { return PASSES == 3 && BN == 128; }
__host__ __device__ constexpr int stage_bytes2(int BN, int PASSES) {
    return stack2(BN, PASSES) ? 2 * A_TILE_BYTES + 128 * BK * 2 + 64 * BK * 2 : (PASSES >= 3 ? 2 : 1) * (A_TILE_BYTES + (BN / 2) * BK * 2);
}
__host__ __device__ constexpr int num_stages2(int BN, int PASSES) {
    return kSmemBudget / stage_bytes2(BN, PASSES) > 8 ? 8 : kSmemBudget / stage_bytes2(BN, PASSES);
}

// Layer chains (round 2).  A layer whose pixel tiles do not fill a whole number of waves (the 32x32 / 40x40 maps: 1.7 - 5.4 waves
// of tile pairs on 74 SM pairs) leaves 10 - 14 % of the machine idle in its last wave, and a dependent launch cannot use that time
// because griddepcontrol.wait blocks until the WHOLE previous grid has finished.  Three mechanisms remove the bubble:
//   * dynamic tile scheduler: the leader's producer warp draws tile tickets from a device counter (p.sched) and hands every ticket
//     to the MMA warp, the epilogue warps and the peer CTA through a 4-slot shared-memory ring (sfull / sempty mbarriers, the peer
//     copy written through DSMEM) - the CTAs that become resident early simply take the first tickets; results do not depend on
//     which CTA computes a tile, so this changes no bit of the output;
//   * per-image completion counters: after the stores of a tile every epilogue thread fences, the eight epilogue warps meet at a
//     named barrier and one thread bumps sig_cnt[image]; the NEXT layer's producer warps wait for dep_cnt[image] == dep_target
//     (acquire at device scope, then fence.proxy.async because TMA reads through the async proxy) instead of griddepcontrol.wait;
//     convolutions never read across images, so a tile may start as soon as its own images are complete;
//   * stores stay behind griddepcontrol.wait (executed by each epilogue thread before its first store): the ping-pong slots make
//     the output of layer i+1 alias the input of layer i, so layer i+1 may READ early but never WRITE before layer i has
//     completed.  Completion of a kernel therefore still implies completion of all its predecessors.
// A CTA of layer i+1 that became resident in the tail of layer i thus runs the main loop of its first tile, drains the accumulator
// chunks into registers, and blocks only at its first store.  All waits are bounded (trap + error word).
constexpr int kSchedSlots = 4;
constexpr int kSchedConsumers = 18;   // MMA warp + 8 epilogue warps of the leader, producer warp + 8 epilogue warps of the peer

template <int BN, int PASSES, bool FP16>
// Register budget: 10 warps land 3 / 3 / 2 / 2 on the four SM sub-partitions of 16384 registers each, so 16384 / (3 x 32) = 170 ->
// 168 registers per thread is the hardware limit for this block shape (a __maxnreg__(200) build compiles spill-free at 197
// registers but fails to launch: "too many resources requested").  With BN = 256 the 128 fp32 partial sums per epilogue thread
// therefore spill 284 bytes to (L1-resident) local memory; the epilogue is off the tensor pipe's critical path.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const __grid_constant__ CUtensorMap map_x_h8, const __grid_constant__ CUtensorMap map_w_l8, const TcParams p) {
    constexpr int STAGES = num_stages2(BN, PASSES);
    constexpr int STAGE_BYTES = stage_bytes2(BN, PASSES);
    constexpr int B_TILE_BYTES = (BN / 2) * BK * 2;
    constexpr int A8_TILE_BYTES = BM * BK, B8_TILE_BYTES = (BN / 2) * BK;   // e4m3 tiles (mode 4): 64-byte rows
    constexpr uint32_t IDESC = make_idesc(BN, FP16, 256);
    constexpr bool STACK = stack2(BN, PASSES);
    constexpr uint32_t IDESC_STACK = make_idesc(2 * BN, FP16, 256);
    constexpr int ACC_COLS = STACK ? 2 * BN : BN;      // TMEM columns per accumulator stage
    constexpr int TMEM_COLS = tmem_cols(ACC_COLS);
    constexpr int Y_BYTES = 128 * BK * 2;              // stacked variant: weight region Y
    static_assert(PASSES != 4 || FP16, "fp8-correction mode uses an fp16 main plane");
    constexpr int COLS = BN / 2;                       // accumulator columns drained by one epilogue warp
    static_assert(STAGES >= 2, "need at least a double-buffered pipeline");
    static_assert((2 * STAGES + 4 + 2 * kSchedSlots) * 8 + kSchedSlots * 4 + 4 <= 256, "barrier area");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* sfull_bar = tempty_bar + 2;              // tile ring: ticket published (one per CTA, arrived by the leader's producer)
    uint64_t* sempty_bar = sfull_bar + kSchedSlots;    // tile ring: ticket consumed (leader CTA only, kSchedConsumers arrivals)
    int* sched_tile = reinterpret_cast<int*>(sempty_bar + kSchedSlots);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_tile + kSchedSlots);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int kblocks = p.k * p.k * p.cin_chunks;

    if (warp == 8 && lane == 0) {
        prefetch_tmap(&map_x_hi); prefetch_tmap(&map_w_hi);
        if (PASSES >= 3) { prefetch_tmap(&map_x_lo); prefetch_tmap(&map_w_lo); }
        if (PASSES == 4) { prefetch_tmap(&map_x_h8); prefetch_tmap(&map_w_l8); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 16); }
        for (int s = 0; s < kSchedSlots; ++s) { mbar_init(&sfull_bar[s], 1); mbar_init(&sempty_bar[s], kSchedConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();          // barriers of BOTH CTAs are initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (!p.dep_cnt) pdl_wait();  // chained layers: the producer waits per image, the epilogue before its first store

    if (warp == 8) {
        // ================================ TMA producer (both CTAs; whole warp, one elected lane issues) ================================
        int stage = 0; uint32_t phase = 0;
        for (int seq = 0;; ++seq) {
            const int slot = seq & (kSchedSlots - 1);
            const uint32_t sph = (uint32_t)(seq / kSchedSlots) & 1u;
            int item;
            if (rank == 0) {     // draw the next ticket and publish it to both CTAs
                mbar_wait(&sempty_bar[slot], sph ^ 1, p.err_flag, 5);
                int v = 0;
                if (lane == 0) {
                    v = p.sched ? atomicAdd(p.sched, 1) : cluster_id + seq * num_clusters;
                    if (v >= p.num_tiles) v = -1;
                    sched_tile[slot] = v;
                    st_shared_cluster_u32(&sched_tile[slot], 1, (uint32_t)v);
                    mbar_arrive(&sfull_bar[slot]);
                    mbar_arrive_cluster_release(&sfull_bar[slot], 1);
                }
                item = __shfl_sync(0xFFFFFFFFu, v, 0);
            } else {
                mbar_wait_cl(&sfull_bar[slot], sph, p.err_flag, 6);
                item = sched_tile[slot];
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&sempty_bar[slot], 0);
            }
            item = __reduce_max_sync(0xFFFFFFFFu, item);   // same value in every lane: keeps the TMA coordinates in uniform registers
            if (item < 0) break;
            const int nt = item % p.n_tiles, mt = 2 * (item / p.n_tiles) + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
            const int w0 = tw * p.TW - p.pad, h0 = th * p.TH - p.pad, b0 = tb * p.TB, n0 = nt * BN + (int)rank * (BN / 2);
            if (p.dep_cnt) {     // the images of this tile are complete in the producing layer (an odd tile count leaves b0 >= B: zero fill)
                const int bend = min(b0 + p.TB, p.B);
                for (int b = b0; b < bend; ++b) counter_wait(p.dep_cnt + b, p.dep_target, p.err_flag, 7);
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            int kcol = 0;
            for (int kh = 0; kh < p.k; ++kh) {
                for (int kw = 0; kw < p.k; ++kw) {
                    for (int cc = 0; cc < p.cin_chunks; ++cc, kcol += BK) {
                        mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
                        if (elect_one()) {
                            uint8_t* st = smem + stage * STAGE_BYTES;
                            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);   // bytes of both CTAs
                            tma_load_4d_2sm(&map_x_hi, st, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                            if (STACK) {   // [A_hi | A_lo | Y | X]; map_w_hi / map_w_lo have 128-row boxes, map_w_l8 = W_hi with 64-row boxes
                                tma_load_4d_2sm(&map_x_lo, st + A_TILE_BYTES, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                                tma_load_2d_2sm(rank == 0 ? &map_w_hi : &map_w_lo, st + 2 * A_TILE_BYTES, &full_bar[stage], kcol, nt * BN);
                                tma_load_2d_2sm(&map_w_l8, st + 2 * A_TILE_BYTES + Y_BYTES, &full_bar[stage], kcol, n0);
                            } else {
                                tma_load_2d_2sm(&map_w_hi, st + (PASSES >= 3 ? 2 : 1) * A_TILE_BYTES, &full_bar[stage], kcol, n0);
                                if (PASSES == 3) {
                                    tma_load_4d_2sm(&map_x_lo, st + A_TILE_BYTES, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                                    tma_load_2d_2sm(&map_w_lo, st + 2 * A_TILE_BYTES + B_TILE_BYTES, &full_bar[stage], kcol, n0);
                                }
                                if (PASSES == 4) {
                                    tma_load_4d_2sm(&map_x_lo, st + A_TILE_BYTES, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                                    tma_load_4d_2sm(&map_x_h8, st + A_TILE_BYTES + A8_TILE_BYTES, &full_bar[stage], cc * BK, w0 + kw, h0 + kh, b0);
                                    tma_load_2d_2sm(&map_w_lo, st + 2 * A_TILE_BYTES + B_TILE_BYTES, &full_bar[stage], kcol, n0);
                                    tma_load_2d_2sm(&map_w_l8, st + 2 * A_TILE_BYTES + B_TILE_BYTES + B8_TILE_BYTES, &full_bar[stage], kcol, n0);
                                }
                            }
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ================================ MMA issuer (leader CTA; whole warp, one elected lane issues) ================================
        if (rank == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc_it = 0;
            for (int seq = 0;; ++seq) {
                const int slot = seq & (kSchedSlots - 1);
                mbar_wait(&sfull_bar[slot], (uint32_t)(seq / kSchedSlots) & 1u, p.err_flag, 6);   // written by this CTA's producer
                int item = sched_tile[slot];
                __syncwarp();
                if (lane == 0) mbar_arrive(&sempty_bar[slot]);
                item = __reduce_max_sync(0xFFFFFFFFu, item);
                if (item < 0) break;
                for (int kb0 = 0; kb0 < kblocks; kb0 += p.chunk_kb, ++acc_it) {
                    const int acc = acc_it & 1;
                    mbar_wait(&tempty_bar[acc], ((acc_it >> 1) & 1) ^ 1, p.err_flag, 2);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
                    const int kb1 = min(kblocks, kb0 + p.chunk_kb);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                            const uint64_t a_hi = make_smem_desc(sa);
                            const uint64_t a_lo = make_smem_desc(sa + A_TILE_BYTES);
                            const uint64_t b_hi = make_smem_desc(sa + (PASSES >= 3 ? 2 : 1) * A_TILE_BYTES);
                            const uint64_t b_lo = make_smem_desc(sa + 2 * A_TILE_BYTES + (STACK ? Y_BYTES : B_TILE_BYTES));
#pragma unroll
                            for (int j = 0; j < BK / UMMA_K; ++j) {
                                const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                                if (STACK) {   // b_hi = region Y ([W_hi ; W_lo] across the pair), b_lo = region X (W_hi halves)
                                    tc_mma_f16_2cta(d_tmem, a_hi + koff, b_hi + koff, IDESC_STACK, (uint32_t)((kb > kb0) | (j != 0)));
                                    tc_mma_f16_2cta(d_tmem, a_lo + koff, b_lo + koff, IDESC, 1u);
                                } else {
                                    tc_mma_f16_2cta(d_tmem, a_hi + koff, b_hi + koff, IDESC, (uint32_t)((kb > kb0) | (j != 0)));
                                    if (PASSES == 3) {
                                        tc_mma_f16_2cta(d_tmem, a_hi + koff, b_lo + koff, IDESC, 1u);
                                        tc_mma_f16_2cta(d_tmem, a_lo + koff, b_hi + koff, IDESC, 1u);
                                    }
                                }
                            }
                            if (PASSES == 4) {
                                const uint64_t a_l8 = make_smem_desc64(sa + A_TILE_BYTES), a_h8 = make_smem_desc64(sa + A_TILE_BYTES + A8_TILE_BYTES);
                                const uint64_t b_h8 = make_smem_desc64(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
                                const uint64_t b_l8 = make_smem_desc64(sa + 2 * A_TILE_BYTES + B_TILE_BYTES + B8_TILE_BYTES);
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint64_t koff = (uint64_t)((j * 32) >> 4);
                                    tc_mma_f8_2cta(d_tmem, a_l8 + koff, b_h8 + koff, IDESC, 1u);
                                    tc_mma_f8_2cta(d_tmem, a_h8 + koff, b_l8 + koff, IDESC, 1u);
                                }
                            }
                            tc_commit_2cta(&empty_bar[stage]);   // frees this smem stage in both CTAs
                            if (kb + 1 == kb1) tc_commit_2cta(&tfull_bar[acc]);   // partial accumulator complete -> epilogues of both CTAs
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ================================ epilogue (8 warps: lane quadrant q, column half ch) ================================
        const int q = warp & 3, ch = warp >> 2;
        const int row = q * 32 + lane;                       // tile row = TMEM lane
        const int w_l = row % p.TW, h_l = (row / p.TW) % p.TH, b_l = row / (p.TW * p.TH);
        int acc_it = 0;
        bool may_store = p.dep_cnt == nullptr;               // chained layers: griddepcontrol.wait before the first store (WAR on the slots)
        for (int seq = 0;; ++seq) {
            const int slot = seq & (kSchedSlots - 1);
            // the peer's copy of the ticket was stored through DSMEM by the leader: acquire at cluster scope there (it invalidates L1)
            if (rank == 0) mbar_wait(&sfull_bar[slot], (uint32_t)(seq / kSchedSlots) & 1u, p.err_flag, 6);
            else mbar_wait_cl(&sfull_bar[slot], (uint32_t)(seq / kSchedSlots) & 1u, p.err_flag, 6);
            const int item = sched_tile[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&sempty_bar[slot], 0);
            if (item < 0) break;
            const int nt = item % p.n_tiles, mt = 2 * (item / p.n_tiles) + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
            const int w = tw * p.TW + w_l, h = th * p.TH + h_l, b = tb * p.TB + b_l, n0 = nt * BN + ch * COLS;
            bool valid = (w < p.W) && (h < p.H) && (b < p.B);
            int64_t pix = ((int64_t)b * p.H + h) * p.W + w;
            if (p.pool) {
                const int par = p.pool == 2 ? 1 : 0;
                valid = valid && ((w & 1) == par) && ((h & 1) == par);
                pix = ((int64_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            }
            float racc[COLS];
            for (int kb0 = 0; kb0 < kblocks; kb0 += p.chunk_kb, ++acc_it) {
                const int acc = acc_it & 1;
                mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, p.err_flag, 4);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + ch * COLS);
#pragma unroll
                for (int c0 = 0; c0 < COLS; c0 += 32) {
                    uint32_t v[32];
                    tc_ld_32x32b_x32(taddr + c0, v);
                    tc_wait_ld();
                    if (kb0 == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) racc[c0 + i] = __uint_as_float(v[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) racc[c0 + i] += __uint_as_float(v[i]);
                    }
                    if (STACK) {   // + the hi*lo products of the stacked half
                        tc_ld_32x32b_x32(taddr + BN + c0, v);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) racc[c0 + i] += __uint_as_float(v[i]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);   // leader's barrier: 16 arrivals per phase
            }
            if (!may_store) { pdl_wait(); may_store = true; }
#pragma unroll
            for (int c0 = 0; c0 < COLS; c0 += 32) epilogue_store32<PASSES, FP16>(p, &racc[c0], pix, n0 + c0, valid);
            if (p.sig_cnt) {     // this CTA's tile is stored: bump the completion counters of its images
                __threadfence();
                named_bar_sync(1, 32 * 8);
                if (threadIdx.x == 0) {
                    const int b0 = tb * p.TB, bend = min(b0 + p.TB, p.B);
                    for (int bb = b0; bb < bend; ++bb) atomicAdd(p.sig_cnt + bb, 1);
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer may still touch this CTA's smem / TMEM
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ 64-channel kernel on a CTA pair
// conv_c64_kernel is bound by shared-memory operand reads (N = 64: 14 KB per 96 tensor cycles = 146 B/clk against 128 B/clk;
// ncu r01e: tensor pipe busy 76 %, computing 56 %) and by its two-deep patch ring (the 144 KB of resident weights leave room
// for two 40 KB stages).  The cta_group::2 version halves the weight bytes per SM and the B rows each SM feeds per instruction:
//   * CTA r keeps, per tap, [W_hi[32 r .. 32 r + 32) ; W_lo[32 r .. 32 r + 32)] (64 rows = 8 KB; 72 KB for the layer), which
//     leaves three 40 KB patch stages;
//   * UMMA 1: A_hi x stacked B, M = 256 (CTA r owns pixel tile 2 * item + r), N = 128: columns [0,32) hi*hi ch 0-31, [32,64) hi*lo
//     ch 0-31 (CTA 0's rows), [64,96) hi*hi ch 32-63, [96,128) hi*lo ch 32-63 (CTA 1's rows);
//   * UMMA 2: A_lo x W_hi, N = 64, B = the first 32 rows of the same weight region in both CTAs, into columns [128,192):
//     lo*hi ch 0-31 | ch 32-63 (a separate accumulator region because the column order differs);
//   * per pair of UMMAs each SM reads 11 KB instead of 14 KB; 8 epilogue warps (lane quadrant x 32-channel half) add the three
//     column groups.
// Barriers as in conv_tc2_kernel.  p.num_tiles counts pixel-tile PAIRS; Cout = 64 * n_tiles: cluster c serves channel group c % n_tiles.
constexpr int C64X2_W_TAP_BYTES = 64 * BK * 2;
constexpr int C64X2_W_BYTES = 9 * C64X2_W_TAP_BYTES;
constexpr int C64X2_A_STAGE_BYTES = 2 * C64_PATCH_BYTES;
constexpr int C64X2_STAGES = 3;
constexpr int C64X2_ACC_COLS = 256;
constexpr int C64X2_OUT_BYTES = 2 * A_TILE_BYTES;        // un-pooled layers: the 128 x 64 output tile (hi, lo) is staged here and leaves by bulk tensor stores
constexpr int C64X2_SMEM = C64X2_W_BYTES + C64X2_STAGES * C64X2_A_STAGE_BYTES + C64X2_OUT_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(C64X2_SMEM <= 227 * 1024, "conv_c64x2_kernel: shared memory budget");

template <bool FP16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
conv_c64x2_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                  const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                  const __grid_constant__ CUtensorMap map_y_hi, const __grid_constant__ CUtensorMap map_y_lo, const TcParams p) {
    constexpr uint32_t IDESC_MAIN = make_idesc(128, FP16, 256);
    constexpr uint32_t IDESC_N64 = make_idesc(64, FP16, 256);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem;
    uint8_t* asm_ = smem + C64X2_W_BYTES;
    uint8_t* osm = asm_ + C64X2_STAGES * C64X2_A_STAGE_BYTES;           // output staging [hi tile | lo tile], 1024-byte aligned
    uint64_t* a_full = reinterpret_cast<uint64_t*>(osm + C64X2_OUT_BYTES);
    uint64_t* a_empty = a_full + C64X2_STAGES;
    uint64_t* tfull_bar = a_empty + C64X2_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_full = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int n0 = (cluster_id % p.n_tiles) * 64;
    const int item0 = cluster_id / p.n_tiles, item_step = num_clusters / p.n_tiles;

    if (warp == 8 && lane == 0) {
        prefetch_tmap(&map_x_hi); prefetch_tmap(&map_w_hi); prefetch_tmap(&map_x_lo); prefetch_tmap(&map_w_lo);
        for (int s = 0; s < C64X2_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 16); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * C64X2_ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (warp != 8) pdl_wait();   // the TMA producer warp waits after it has issued the (static) weight loads

    if (warp == 8) {
        // ================================ TMA producer (both CTAs; whole warp, one elected lane issues) ================================
        if (elect_one()) {
            if (rank == 0) mbar_expect_tx(w_full, 2 * C64X2_W_BYTES);
            for (int t = 0; t < 9; ++t) {
                tma_load_2d_2sm(&map_w_hi, wsm + t * C64X2_W_TAP_BYTES, w_full, t * BK, n0 + (int)rank * 32);
                tma_load_2d_2sm(&map_w_lo, wsm + t * C64X2_W_TAP_BYTES + 32 * BK * 2, w_full, t * BK, n0 + (int)rank * 32);
            }
        }
        __syncwarp();
        pdl_wait();
        int stage = 0; uint32_t phase = 0;
        for (int item = item0; item < p.num_tiles; item += item_step) {
            const int mt = 2 * item + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
            const int w0 = tw * C64_TW - 1, h0 = th * C64_TH - 1;
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&a_empty[stage], phase ^ 1, p.err_flag, 1);
                if (elect_one()) {
                    uint8_t* st = asm_ + stage * C64X2_A_STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(&a_full[stage], 2 * C64X2_A_STAGE_BYTES);
                    tma_load_4d_2sm(&map_x_hi, st, &a_full[stage], 0, w0 + kw, h0, tb);   // tb >= B (odd tile count): zero fill
                    tma_load_4d_2sm(&map_x_lo, st + C64_PATCH_BYTES, &a_full[stage], 0, w0 + kw, h0, tb);
                }
                __syncwarp();
                if (++stage == C64X2_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 9) {
        // ================================ MMA issuer (leader CTA; whole warp, one elected lane issues) ================================
        if (rank == 0) {
            mbar_wait(w_full, 0, p.err_flag, 5);
            tc_fence_after();
            const uint32_t wb = smem_u32(wsm);
            int stage = 0; uint32_t phase = 0;
            int acc_it = 0;
            for (int item = item0; item < p.num_tiles; item += item_step, ++acc_it) {
                const int acc = acc_it & 1;
                mbar_wait(&tempty_bar[acc], ((acc_it >> 1) & 1) ^ 1, p.err_flag, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C64X2_ACC_COLS);
                for (int kw = 0; kw < 3; ++kw) {
                    mbar_wait(&a_full[stage], phase, p.err_flag, 3);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(asm_ + stage * C64X2_A_STAGE_BYTES);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint64_t a_hi = make_smem_desc(sa + kh * C64_ROW_BYTES);
                            const uint64_t a_lo = make_smem_desc(sa + C64_PATCH_BYTES + kh * C64_ROW_BYTES);
                            const uint64_t b = make_smem_desc(wb + (kh * 3 + kw) * C64X2_W_TAP_BYTES);
#pragma unroll
                            for (int j = 0; j < BK / UMMA_K; ++j) {
                                const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                                const uint32_t accum = (uint32_t)((kw | kh | j) != 0);
                                tc_mma_f16_2cta(d_tmem, a_hi + koff, b + koff, IDESC_MAIN, accum);
                                tc_mma_f16_2cta(d_tmem + 128, a_lo + koff, b + koff, IDESC_N64, accum);
                            }
                        }
                        tc_commit_2cta(&a_empty[stage]);
                        if (kw == 2) tc_commit_2cta(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == C64X2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue (8 warps: lane quadrant q, 32-channel half ch) ================================
        const int q = warp & 3, ch = warp >> 2;
        const int row = q * 32 + lane;
        const int w_l = row % C64_TW, h_l = row / C64_TW;
        int acc_it = 0;
        for (int item = item0; item < p.num_tiles; item += item_step, ++acc_it) {
            const int mt = 2 * item + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, b = mt / (p.tiles_w * p.tiles_h);
            const int w = tw * C64_TW + w_l, h = th * C64_TH + h_l;
            bool valid = (w < p.W) && (h < p.H) && (b < p.B);
            int64_t pix = ((int64_t)b * p.H + h) * p.W + w;
            if (p.pool) {
                const int par = p.pool == 2 ? 1 : 0;
                valid = valid && ((w & 1) == par) && ((h & 1) == par);
                pix = ((int64_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            }
            const int acc = acc_it & 1;
            mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, p.err_flag, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C64X2_ACC_COLS);
            uint32_t v0[32], v1[32], v2[32];
            tc_ld_32x32b_x32(taddr + 64 * ch, v0);          // hi*hi
            tc_ld_32x32b_x32(taddr + 64 * ch + 32, v1);     // hi*lo
            tc_ld_32x32b_x32(taddr + 128 + 32 * ch, v2);    // lo*hi
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
            float racc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) racc[i] = (__uint_as_float(v0[i]) + __uint_as_float(v1[i])) + __uint_as_float(v2[i]);
            if (!p.tma_out) {
                epilogue_store32<3, FP16>(p, racc, pix, n0 + 32 * ch, valid);
                continue;
            }
            // Un-pooled layer (conv2_1): 8 bytes per output value leave the SM.  Direct 16-byte st.global touches 32 different lines
            // per warp instruction (the thread <-> pixel mapping of TMEM) and made this layer LSU bound (ncu r02d: tensor pipe 50 %
            // with either kernel family); instead the tile is staged in the SWIZZLE_128B layout and written by one bulk tensor store
            // per plane (full 128-byte lines, partial tiles clipped by TMA), exactly as in conv_c3_tma_kernel.
            uint32_t hi[16], lo[16];
            const float2* bias2 = reinterpret_cast<const float2*>(p.bias + n0 + 32 * ch);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 bv = __ldg(bias2 + i);
                float f0 = racc[2 * i] + bv.x, f1 = racc[2 * i + 1] + bv.y;
                if (p.leaky) { f0 = fmaxf(f0, kNegSlope * f0); f1 = fmaxf(f1, kNegSlope * f1); }
                hi[i] = pack_hi2<FP16>(f0, f1);
                const float2 r = unpack2<FP16>(hi[i]);
                lo[i] = pack_hi2<FP16>(f0 - r.x, f1 - r.y);
            }
            if (warp == 0) { if (elect_one()) tma_store_wait_read(); __syncwarp(); }   // the previous tile's stores have read the staging buffer
            named_bar_sync(2, 256);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                sts_v4(smem_u32(osm) + sw128_chunk(row, 4 * ch + c), hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                sts_v4(smem_u32(osm) + A_TILE_BYTES + sw128_chunk(row, 4 * ch + c), lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(2, 256);
            if (warp == 0) {
                if (elect_one() && b < p.B) {                   // (an odd tile count leaves the pair's second tile outside the batch)
                    tma_store_4d(&map_y_hi, osm, n0, tw * C64_TW, th * C64_TH, b);
                    tma_store_4d(&map_y_lo, osm + A_TILE_BYTES, n0, tw * C64_TW, th * C64_TH, b);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (p.tma_out && warp == 0) { if (elect_one()) tma_store_wait_all(); __syncwarp(); }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * C64X2_ACC_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ conv1_1 fused into conv1_2 (CTA pair)
// conv1_1 (3 -> 64 channels) writes 8 bytes per output value that conv1_2 reads straight back: 1.3 GB written and 1.3 GB read per
// 32-image step, 0.3 ms of HBM-bound launches (conv_c3_tma_kernel).  This kernel is conv_c64x2_kernel for conv1_2 whose activation
// patches are not fetched by TMA but COMPUTED in place, by a second (tiny) tensor-core GEMM:
//   * a pixel tile of conv1_2 (16 x 8) needs conv1_1's output on the 18 x 10 patch around it = 180 patch pixels; they are the rows of
//     the conv1_1 GEMM (two blocks of 128 rows per CTA), K = 27 -> 32, N = 64;
//   * the bias of conv1_1 is K index 27 of the GEMM (a constant-1 column of A times a bias row of W, both hi / lo split), so the
//     mid-epilogue has no bias loads or adds;
//   * four BUILDER warps stage the 20 x 12 x 3 fp32 image patch (the loads of the next tile are in flight while this one is
//     converted) and write the im2col rows [A_hi 64 B | A_lo 64 B] into ONE 16 KB operand buffer, block after block;
//   * the MMA warp of the leader interleaves the conv1_1 instructions of tile t + 1 (cta_group::2, M = 256 = this block of both CTAs,
//     N = 64, three passes into one accumulator: 12 UMMAs per tile) with the kw groups of conv1_2's tile t - block 0 before the first
//     group, block 1 after it; their accumulators use the 2 x 64 TMEM columns the conv1_2 accumulator stages leave free ([192,256) and
//     [448,512));
//   * eight MID-epilogue warps: TMEM -> leaky ReLU, ZERO outside the image (conv1_2's 'SAME' padding applies to conv1_1's OUTPUT) ->
//     hi / lo split kept in registers -> written into the three kw-shifted patch stages of the conv_c64 scheme (canonical SWIZZLE_128B
//     rows, st.shared + fence.proxy.async + mbarrier arrive) as the stages are released by conv1_2's MMAs of the previous tile;
//   * eight FINAL-epilogue warps: bias, leaky ReLU, 2 x 2 max-pool, split, store - as in conv_c64x2_kernel.  Mid- and final epilogue
//     are separate warps because each is a ~3000-cycle serial chain per tile against 3840 tensor cycles per tile: on the same warps
//     (first version of this kernel) the tile period was their SUM (profiles/r02g_c1f_sampling.md).
// Barriers (leader's instance counts both CTAs): a1_full 8 builder warps -> MMA, a1_empty commit -> builders, t1_full commit -> mid-
// epilogue, t1_empty 16 warps -> MMA (values are in registers: the columns may be overwritten), a_full[kw] 16 warps -> MMA, a_empty[kw]
// commit -> mid-epilogue, tfull / tempty as before.  Stage index = kw, parity = tile iteration.  All waits bounded.
constexpr int C1F_THREADS = 32 * 21;                       // warps 0-7 final epilogue, 8-15 mid-epilogue, 16-19 builders, 20 MMA issuer
constexpr int C1F_BUILDERS = 128;                          // (two builder warps are not enough: they become the critical path, ncu r02i)
constexpr int C1F_PPX = C64_TW + 2, C1F_PPY = C64_TH + 2;  // conv1_1 output patch 18 x 10
constexpr int C1F_PP = C1F_PPX * C1F_PPY;                  // 180 patch pixels = GEMM rows (2 blocks of 128)
constexpr int C1F_IPX = C1F_PPX + 2, C1F_IPY = C1F_PPY + 2;
constexpr int C1F_IMG_FLOATS = C1F_IPY * C1F_IPX * 3;      // 720: image patch 12 x 20 x 3
constexpr int C1F_A1_BYTES = A_TILE_BYTES;                 // 128 rows x [A_hi 64 B | A_lo 64 B]
constexpr int C1F_W1_BYTES = 64 * BK * 2;                  // 64 rows x 128 B (first 64 B used): W_hi[32 r ..) rows 0-31, W_lo rows 32-63
constexpr int C1F_SMEM = C64X2_W_BYTES + C64X2_STAGES * C64X2_A_STAGE_BYTES + C1F_A1_BYTES + C1F_W1_BYTES + 2 * C1F_IMG_FLOATS * 4 +
                         1024 /*align slack*/ + 256 /*barriers*/;
static_assert(C1F_SMEM <= 227 * 1024, "conv_c1f_kernel: shared memory budget");

template <bool FP16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C1F_THREADS, 1)
conv_c1f_kernel(const float* __restrict__ x, const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                const TcParams p) {
    constexpr uint32_t IDESC_MAIN = make_idesc(128, FP16, 256);
    constexpr uint32_t IDESC_N64 = make_idesc(64, FP16, 256);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem;                                              // conv1_2 weights, resident (as conv_c64x2_kernel)
    uint8_t* asm_ = smem + C64X2_W_BYTES;                             // three kw patch stages [hi 20 KB | lo 20 KB]
    uint8_t* a1sm = asm_ + C64X2_STAGES * C64X2_A_STAGE_BYTES;        // conv1_1 A operand, one 128-row block
    uint8_t* w1sm = a1sm + C1F_A1_BYTES;                              // conv1_1 B operand halves of this CTA
    float* imgsm = reinterpret_cast<float*>(w1sm + C1F_W1_BYTES);     // [2][720] image patch
    uint64_t* a_full = reinterpret_cast<uint64_t*>(imgsm + 2 * C1F_IMG_FLOATS);
    uint64_t* a_empty = a_full + 3;
    uint64_t* tfull_bar = a_empty + 3;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* w_full = tempty_bar + 2;
    uint64_t* a1_full = w_full + 1;
    uint64_t* a1_empty = a1_full + 1;
    uint64_t* t1_full = a1_empty + 1;
    uint64_t* t1_empty = t1_full + 1;
    uint64_t* w1_full = t1_empty + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w1_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 20 && lane == 0) {
        prefetch_tmap(&map_w_hi); prefetch_tmap(&map_w_lo);
        for (int s = 0; s < 3; ++s) { mbar_init(&a_full[s], 16); mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 16); }
        mbar_init(w_full, 1);
        mbar_init(a1_full, 8); mbar_init(a1_empty, 1); mbar_init(t1_full, 1); mbar_init(t1_empty, 16); mbar_init(w1_full, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 19) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * C64X2_ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    if (warp != 20) pdl_wait();   // the MMA warp waits after it has issued the (static) weight loads

    if (warp >= 16 && warp < 20) {
        // ================================ builders: conv1_1 weights once, then the im2col rows of every tile ================================
        const int t = threadIdx.x - 512;                   // 0 .. 127
        if (t < 64) {                                      // row t: plane t >> 5 (0 hi, 1 lo) of output channel 32 rank + (t & 31)
            const int co = 32 * (int)rank + (t & 31);
            uint32_t pk[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                float v0 = 0.f, v1 = 0.f;
                if (2 * k2 < 27) v0 = __ldg(p.w1 + (2 * k2) * 64 + co);
                if (2 * k2 + 1 < 27) v1 = __ldg(p.w1 + (2 * k2 + 1) * 64 + co);
                if (2 * k2 + 1 == 27) v1 = __ldg(p.bias1 + co);      // K index 27: the bias, multiplied by the constant-1 column of A
                const uint32_t h = pack_hi2<FP16>(v0, v1);
                if (t < 32) pk[k2] = h;
                else { const float2 r = unpack2<FP16>(h); pk[k2] = pack_hi2<FP16>(v0 - r.x, v1 - r.y); }
            }
            const uint32_t w1a = smem_u32(w1sm);
#pragma unroll
            for (int c = 0; c < 4; ++c) sts_v4(w1a + sw128_chunk(t, c), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(w1_full, 0);
        constexpr int PRE = (C1F_IMG_FLOATS + C1F_BUILDERS - 1) / C1F_BUILDERS;  // 6
        float pre[PRE];
        auto load_img = [&](int item) {
            const int mt = 2 * item + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, b = mt / (p.tiles_w * p.tiles_h);
            const int x0 = tw * C64_TW - 2, y0 = th * C64_TH - 2;
            const float* xb = x + (int64_t)b * p.H * p.W * 3;
#pragma unroll
            for (int j = 0; j < PRE; ++j) {
                const int i = t + j * C1F_BUILDERS;
                const int r = i / (C1F_IPX * 3), rem = i - r * (C1F_IPX * 3);
                const int gy = y0 + r, gx = x0 + rem / 3;
                float v = 0.f;
                if (i < C1F_IMG_FLOATS && b < p.B && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(xb + ((int64_t)gy * p.W + x0) * 3 + rem);
                pre[j] = v;
            }
        };
        const uint32_t a1a = smem_u32(a1sm);
        int fills = 0, it = 0;
        if (cluster_id < p.num_tiles) load_img(cluster_id);
        for (int item = cluster_id; item < p.num_tiles; item += num_clusters, ++it) {
            const uint32_t pb = smem_u32(imgsm + (it & 1) * C1F_IMG_FLOATS);
#pragma unroll
            for (int j = 0; j < PRE; ++j)
                if (t + j * C1F_BUILDERS < C1F_IMG_FLOATS) sts_f32(pb + 4u * (uint32_t)(t + j * C1F_BUILDERS), pre[j]);
            named_bar_sync(1, C1F_BUILDERS);
            if (item + num_clusters < p.num_tiles) load_img(item + num_clusters);
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk, ++fills) {
                const int pp = blk * 128 + t;              // patch pixel of this row
                const int py = pp / C1F_PPX, px = pp - py * C1F_PPX;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    float v0 = 0.f, v1 = 0.f;
                    if (pp < C1F_PP) {
                        if (2 * k2 < 27) { const int k = 2 * k2; v0 = lds_f32(pb + 4u * (uint32_t)((py + k / 9) * (C1F_IPX * 3) + px * 3 + (k % 9))); }
                        if (2 * k2 + 1 < 27) { const int k = 2 * k2 + 1; v1 = lds_f32(pb + 4u * (uint32_t)((py + k / 9) * (C1F_IPX * 3) + px * 3 + (k % 9))); }
                        if (2 * k2 + 1 == 27) v1 = 1.0f;             // constant-1 column: the bias comes out of the GEMM
                    }
                    hi[k2] = pack_hi2<FP16>(v0, v1);
                    const float2 r = unpack2<FP16>(hi[k2]);
                    lo[k2] = pack_hi2<FP16>(v0 - r.x, v1 - r.y);
                }
                mbar_wait(a1_empty, ((uint32_t)fills & 1u) ^ 1u, p.err_flag, 11);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    sts_v4(a1a + sw128_chunk(t, c), hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    sts_v4(a1a + sw128_chunk(t, 4 + c), lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(a1_full, 0);
            }
        }
    } else if (warp == 20) {
        // ================================ conv1_2 weights (resident, loaded once) + MMA issuer (leader CTA) ================================
        if (elect_one()) {
            if (rank == 0) mbar_expect_tx(w_full, 2 * C64X2_W_BYTES);
            for (int t = 0; t < 9; ++t) {
                tma_load_2d_2sm(&map_w_hi, wsm + t * C64X2_W_TAP_BYTES, w_full, t * BK, (int)rank * 32);
                tma_load_2d_2sm(&map_w_lo, wsm + t * C64X2_W_TAP_BYTES + 32 * BK * 2, w_full, t * BK, (int)rank * 32);
            }
        }
        __syncwarp();
        pdl_wait();
        if (rank == 0) {
            mbar_wait(w_full, 0, p.err_flag, 5);
            mbar_wait(w1_full, 0, p.err_flag, 12);
            tc_fence_after();
            const uint32_t wb = smem_u32(wsm);
            const uint64_t a1_hi = make_smem_desc(smem_u32(a1sm));
            const uint64_t a1_lo = a1_hi + (uint64_t)(64 >> 4);                     // bytes [64,128) of every row
            const uint64_t b1_hi = make_smem_desc(smem_u32(w1sm));
            const uint64_t b1_lo = make_smem_desc(smem_u32(w1sm) + 32 * 128);       // rows 32-63
            int fills = 0;
            // conv1_1 of block blk of the NEXT tile: 6 UMMAs (2 K steps x hi*hi, hi*lo, lo*hi) into columns [blk * 256 + 192, + 64)
            auto conv1_block = [&](int blk, bool last) {
                mbar_wait(a1_full, (uint32_t)fills & 1u, p.err_flag, 13);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + (uint32_t)(blk * C64X2_ACC_COLS + 192);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                        tc_mma_f16_2cta(d, a1_hi + koff, b1_hi + koff, IDESC_N64, (uint32_t)(j != 0));
                        tc_mma_f16_2cta(d, a1_hi + koff, b1_lo + koff, IDESC_N64, 1u);
                        tc_mma_f16_2cta(d, a1_lo + koff, b1_hi + koff, IDESC_N64, 1u);
                    }
                    tc_commit_2cta(a1_empty);
                    if (last) tc_commit_2cta(t1_full);
                }
                __syncwarp();
                ++fills;
            };
            int it = 0;
            if (cluster_id < p.num_tiles) { conv1_block(0, false); conv1_block(1, true); }
            for (int item = cluster_id; item < p.num_tiles; item += num_clusters, ++it) {
                const int acc = it & 1;
                const bool more = item + num_clusters < p.num_tiles;
                // conv1_1 of the NEXT tile goes in as early as possible - block 0 before this tile's first conv1_2 group, block 1 right
                // after it - so that its accumulators are complete one kw group (1150 tensor cycles) into the tile and the mid-epilogue has
                // the remaining two groups to convert them.  t1_empty: the mid-epilogue of THIS tile holds its values in registers.
                if (more) {
                    mbar_wait(t1_empty, (uint32_t)it & 1u, p.err_flag, 14);
                    tc_fence_after();
                    conv1_block(0, false);
                }
                mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1, p.err_flag, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * C64X2_ACC_COLS);
                for (int kw = 0; kw < 3; ++kw) {
                    mbar_wait(&a_full[kw], (uint32_t)it & 1u, p.err_flag, 3);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = smem_u32(asm_ + kw * C64X2_A_STAGE_BYTES);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint64_t a_hi = make_smem_desc(sa + kh * C64_ROW_BYTES);
                            const uint64_t a_lo = make_smem_desc(sa + C64_PATCH_BYTES + kh * C64_ROW_BYTES);
                            const uint64_t b = make_smem_desc(wb + (kh * 3 + kw) * C64X2_W_TAP_BYTES);
#pragma unroll
                            for (int j = 0; j < BK / UMMA_K; ++j) {
                                const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                                const uint32_t accum = (uint32_t)((kw | kh | j) != 0);
                                tc_mma_f16_2cta(d_tmem, a_hi + koff, b + koff, IDESC_MAIN, accum);
                                tc_mma_f16_2cta(d_tmem + 128, a_lo + koff, b + koff, IDESC_N64, accum);
                            }
                        }
                        tc_commit_2cta(&a_empty[kw]);
                        if (kw == 2) tc_commit_2cta(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (more && kw == 0) conv1_block(1, true);
                }
            }
        }
    } else if (warp >= 8) {
        // ================================ mid-epilogue (8 warps: lane quadrant q, 32-channel half ch) ================================
        const int q = warp & 3, ch = (warp - 8) >> 2;
        const int row = q * 32 + lane;
        int ppy[2], ppx[2]; bool in_patch[2];
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            const int pp = blk * 128 + row;
            ppy[blk] = pp / C1F_PPX; ppx[blk] = pp - ppy[blk] * C1F_PPX;
            in_patch[blk] = pp < C1F_PP;
        }
        int it = 0;
        for (int item = cluster_id; item < p.num_tiles; item += num_clusters, ++it) {
            const int mt = 2 * item + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, b = mt / (p.tiles_w * p.tiles_h);
            mbar_wait(t1_full, (uint32_t)it & 1u, p.err_flag, 15);
            tc_fence_after();
            uint32_t hi[2][16], lo[2][16];
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int gy = th * C64_TH - 1 + ppy[blk], gx = tw * C64_TW - 1 + ppx[blk];
                const bool inside = in_patch[blk] && b < p.B && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
#pragma unroll
                for (int h = 0; h < 2; ++h) {              // 16 channels at a time (register budget of the 19-warp block)
                    uint32_t v[16];
                    tc_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * C64X2_ACC_COLS + 192 + 32 * ch + 16 * h), v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float f0 = __uint_as_float(v[2 * i]), f1 = __uint_as_float(v[2 * i + 1]);   // bias included (K index 27)
                        if (p.leaky1) { f0 = fmaxf(f0, kNegSlope * f0); f1 = fmaxf(f1, kNegSlope * f1); }
                        if (!inside) { f0 = 0.f; f1 = 0.f; }         // conv1_2's zero padding, pixels of no image, rows beyond the patch
                        hi[blk][8 * h + i] = pack_hi2<FP16>(f0, f1);
                        const float2 r = unpack2<FP16>(hi[blk][8 * h + i]);
                        lo[blk][8 * h + i] = pack_hi2<FP16>(f0 - r.x, f1 - r.y);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(t1_empty, 0);
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&a_empty[kw], ((uint32_t)it & 1u) ^ 1u, p.err_flag, 1);
                const uint32_t st = smem_u32(asm_ + kw * C64X2_A_STAGE_BYTES);
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    const int j = ppx[blk] - kw;
                    if (in_patch[blk] && j >= 0 && j < C64_TW) {
                        const int r = ppy[blk] * C64_TW + j;       // row of the {64 ch, 16 px, 10 rows} patch
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            sts_v4(st + sw128_chunk(r, 4 * ch + c), hi[blk][4 * c], hi[blk][4 * c + 1], hi[blk][4 * c + 2], hi[blk][4 * c + 3]);
                            sts_v4(st + C64_PATCH_BYTES + sw128_chunk(r, 4 * ch + c), lo[blk][4 * c], lo[blk][4 * c + 1], lo[blk][4 * c + 2], lo[blk][4 * c + 3]);
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&a_full[kw], 0);
            }
        }
    } else {
        // ================================ final epilogue (8 warps: lane quadrant q, 32-channel half ch), as conv_c64x2_kernel ================================
        const int q = warp & 3, ch = warp >> 2;
        const int row = q * 32 + lane;
        const int w_l = row % C64_TW, h_l = row / C64_TW;
        int it = 0;
        for (int item = cluster_id; item < p.num_tiles; item += num_clusters, ++it) {
            const int mt = 2 * item + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, b = mt / (p.tiles_w * p.tiles_h);
            const int w = tw * C64_TW + w_l, h = th * C64_TH + h_l;
            bool valid = (w < p.W) && (h < p.H) && (b < p.B);
            int64_t pix = ((int64_t)b * p.H + h) * p.W + w;
            if (p.pool) {
                const int par = p.pool == 2 ? 1 : 0;
                valid = valid && ((w & 1) == par) && ((h & 1) == par);
                pix = ((int64_t)b * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            }
            const int acc = it & 1;
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1, p.err_flag, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C64X2_ACC_COLS);
            float racc[32];
            {   // three column groups summed one after the other (register budget), in the order of conv_c64x2_kernel: (hh + hl) + lh
                uint32_t v[32];
                tc_ld_32x32b_x32(taddr + 64 * ch, v);           // hi*hi
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) racc[i] = __uint_as_float(v[i]);
                tc_ld_32x32b_x32(taddr + 64 * ch + 32, v);      // hi*lo
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) racc[i] += __uint_as_float(v[i]);
                tc_ld_32x32b_x32(taddr + 128 + 32 * ch, v);     // lo*hi
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) racc[i] += __uint_as_float(v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
            epilogue_store32<3, FP16>(p, racc, pix, 32 * ch, valid);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 19) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * C64X2_ACC_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ FC stacks as ONE kernel
// PosePrior (2050 -> 512 -> 512 -> 63, optional 30-wide bottleneck) and ViewpointNet (4098 -> 256 -> 128 -> 3) fully connected stacks
// (nets/ColorHandPose3DNetwork.py:262-267,297-308; nets/PosePriorNetwork.py:113-116) followed by Rodrigues / flip / rotate
// (:239-247,311-361) in a single launch instead of 5 GEMM launches + 2 split-K FC kernels + the rotation kernel.
//  * A fully connected layer is the 1x1 case of the implicit GEMM above: M = 128 batch rows (TMA zero-fills rows >= B), N = 64 output
//    features per tile, K = in_features in blocks of 64, 3-pass N-stacked UMMAs (hi*hi | hi*lo in one instruction, lo*hi in a second).
//  * One CLUSTER of 8 CTAs per chain: CTA r of the cluster owns the N tiles r, r + 8, ... of every layer, so the 4.2 MB weight matrix of
//    the first layer streams through 8 SMs.  Hidden activations go through global memory (L2-resident, <= 128 KB) as split planes;
//    between layers every thread fences (generic -> async proxy, the next layer reads through TMA) and the cluster synchronises.
//  * Shared-memory ring, TMEM accumulator double-buffering and barrier phases simply continue across layers (a layer is a tile loop).
//  * Two chains = two clusters in the same grid.  The cluster that finishes LAST (atomic ticket at device scope) applies the
//    Rodrigues / flip / rotate epilogue to the canonical coordinates and the view-point vector of both chains: no extra launch,
//    no waiting.
constexpr int kFcMaxLayers = 4;
constexpr int kFcCluster = 8;
struct FcLayer {
    CUtensorMap map_x_hi, map_x_lo, map_w_hi, map_w_lo;
    TcParams p;            // epilogue parameters (bias, outputs, n_valid, leaky); B / geometry fields unused
    int kblocks, m_tiles, n_tiles, pad_;
};
struct FcChain { FcLayer layer[kFcMaxLayers]; int num_layers; int pad_[3]; };
struct FcChainParams {
    FcChain chain[2];
    int num_chains, B;
    const float* can; const float* uxyz; const float* hand_side;   // rotate epilogue (num_chains == 2)
    float* rot; float* out;
    unsigned int* counter;
    int* err_flag;
};

__device__ __forceinline__ void fc_layer_sync() {
    __threadfence();                                        // the layer's outputs: visible at device scope ...
    asm volatile("fence.proxy.async;" ::: "memory");        // ... and to the async proxy (the next layer's TMA loads)
    cluster_sync_all();
}

template <bool FP16>
__global__ void __cluster_dims__(kFcCluster, 1, 1) __launch_bounds__(kThreads, 1)
fc_chain_kernel(const __grid_constant__ FcChainParams P) {
    constexpr int BN = 64, PASSES = 3;
    constexpr int STAGES = num_stages(BN, PASSES);
    constexpr int STAGE_BYTES = stage_bytes(BN, PASSES);
    constexpr int B_TILE_BYTES = BN * BK * 2;
    constexpr uint32_t IDESC = make_idesc(BN, FP16);
    constexpr uint32_t IDESC_STACK = make_idesc(2 * BN, FP16);
    constexpr int ACC_COLS = 2 * BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    constexpr int kChunk = 9;                              // K blocks per TMEM partial sum (as the convolution kernels)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const FcChain& ch = P.chain[blockIdx.x / kFcCluster];

    if (warp == 4 && lane == 0) {
        for (int l = 0; l < ch.num_layers; ++l) {
            prefetch_tmap(&ch.layer[l].map_x_hi); prefetch_tmap(&ch.layer[l].map_x_lo);
            prefetch_tmap(&ch.layer[l].map_w_hi); prefetch_tmap(&ch.layer[l].map_w_lo);
        }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kNumEpilogueWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    // Every warp of every CTA of the cluster runs the SAME layer loop and reaches fc_layer_sync() exactly once per layer.
    int stage = 0; uint32_t phase = 0; int acc_it = 0;
    for (int l = 0; l < ch.num_layers; ++l) {
        const FcLayer& L = ch.layer[l];
        const int items = L.m_tiles * L.n_tiles;
        if (warp == 4) {
            // ================================ TMA producer (whole warp, one elected lane issues) ================================
            for (int item = rank; item < items; item += kFcCluster) {
                const int nt = item % L.n_tiles, mt = item / L.n_tiles;
                for (int kb = 0; kb < L.kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, P.err_flag, 1);
                    if (elect_one()) {
                        uint8_t* st = smem + stage * STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_4d(&L.map_x_hi, st, &full_bar[stage], kb * BK, 0, 0, mt * BM);
                        tma_load_4d(&L.map_x_lo, st + A_TILE_BYTES, &full_bar[stage], kb * BK, 0, 0, mt * BM);
                        tma_load_2d(&L.map_w_hi, st + 2 * A_TILE_BYTES, &full_bar[stage], kb * BK, nt * BN);
                        tma_load_2d(&L.map_w_lo, st + 2 * A_TILE_BYTES + B_TILE_BYTES, &full_bar[stage], kb * BK, nt * BN);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 5) {
            // ================================ MMA issuer (whole warp, one elected lane issues) ================================
            for (int item = rank; item < items; item += kFcCluster) {
                for (int kb0 = 0; kb0 < L.kblocks; kb0 += kChunk, ++acc_it) {
                    const int acc = acc_it & 1;
                    mbar_wait(&tempty_bar[acc], ((acc_it >> 1) & 1) ^ 1, P.err_flag, 2);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
                    const int kb1 = min(L.kblocks, kb0 + kChunk);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&full_bar[stage], phase, P.err_flag, 3);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                            const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + A_TILE_BYTES);
                            const uint64_t b_hi = make_smem_desc(sa + 2 * A_TILE_BYTES);
#pragma unroll
                            for (int j = 0; j < BK / UMMA_K; ++j) {
                                const uint64_t koff = (uint64_t)((j * UMMA_K * 2) >> 4);
                                tc_mma_f16(d_tmem, a_hi + koff, b_hi + koff, IDESC_STACK, (uint32_t)((kb > kb0) | (j != 0)));
                                tc_mma_f16(d_tmem, a_lo + koff, b_hi + koff, IDESC, 1u);
                            }
                            tc_commit(&empty_bar[stage]);
                            if (kb + 1 == kb1) tc_commit(&tfull_bar[acc]);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else {
            // ================================ epilogue (warps 0-3 <-> TMEM lanes 32w..32w+31 = batch rows) ================================
            for (int item = rank; item < items; item += kFcCluster) {
                const int nt = item % L.n_tiles, mt = item / L.n_tiles;
                const int row = mt * BM + (int)threadIdx.x;
                const bool valid = row < P.B;
                float racc[BN];
                for (int kb0 = 0; kb0 < L.kblocks; kb0 += kChunk, ++acc_it) {
                    const int acc = acc_it & 1;
                    mbar_wait(&tfull_bar[acc], (acc_it >> 1) & 1, P.err_flag, 4);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * ACC_COLS);
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        uint32_t v[32];
                        tc_ld_32x32b_x32(taddr + c0, v);
                        tc_wait_ld();
                        if (kb0 == 0) {
#pragma unroll
                            for (int q = 0; q < 32; ++q) racc[c0 + q] = __uint_as_float(v[q]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; ++q) racc[c0 + q] += __uint_as_float(v[q]);
                        }
                        tc_ld_32x32b_x32(taddr + BN + c0, v);      // + the hi*lo products of the stacked half
                        tc_wait_ld();
#pragma unroll
                        for (int q = 0; q < 32; ++q) racc[c0 + q] += __uint_as_float(v[q]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) epilogue_store32<PASSES, FP16>(L.p, &racc[c0], (int64_t)row, nt * BN + c0, valid);
            }
        }
        fc_layer_sync();
    }

    // ---- last cluster to finish: Rodrigues + right-hand flip + rotation of the canonical coordinates (both chains' outputs)
    if (P.num_chains == 2 && rank == 0) {
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(P.counter, 1u) == 1u;
            if (s_last) *P.counter = 0u;                    // ready for the next launch (launches are stream ordered)
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int b = warp; b < P.B; b += kThreads / 32) {
                float R[9];
                if (lane == 0) rodrigues_rot_mat(__ldcg(P.uxyz + 3 * b), __ldcg(P.uxyz + 3 * b + 1), __ldcg(P.uxyz + 3 * b + 2), R);
#pragma unroll
                for (int i = 0; i < 9; ++i) R[i] = __shfl_sync(0xFFFFFFFFu, R[i], 0);
                if (P.rot && lane < 9) P.rot[9 * b + lane] = R[lane];
                const bool right = P.hand_side[2 * b + 1] > P.hand_side[2 * b];
                float cb[63 / 32 + 1];
                (void)cb;
                for (int i = lane; i < 63; i += 32) {
                    float c3[3];
                    const int kp = i / 3;
                    c3[0] = __ldcg(P.can + 63 * b + 3 * kp); c3[1] = __ldcg(P.can + 63 * b + 3 * kp + 1); c3[2] = __ldcg(P.can + 63 * b + 3 * kp + 2);
                    const int j = i - kp * 3;
                    const float cz = right ? -c3[2] : c3[2];
                    P.out[63 * b + i] = c3[0] * R[j] + c3[1] * R[3 + j] + cz * R[6 + j];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    cluster_sync_all();          // no CTA of the cluster exits while a peer may still be inside a cluster barrier
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
        set_error("cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// es = element size in bytes: 2 (bf16 / fp16 planes, 128-byte rows, SWIZZLE_128B) or 1 (e4m3 planes, 64-byte rows, SWIZZLE_64B)
bool encode_act_map(CUtensorMap* m, const void* base, int C_total, int C_used, int W, int H, int B, int TW, int TH, int TB, int es = 2) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C_used, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C_total * es, (cuuint64_t)W * C_total * es, (cuuint64_t)H * W * C_total * es};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, es == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return false; }
    return true;
}

bool encode_w_map(CUtensorMap* m, const void* base, int Ktot, int Cout_pad, int BN, int es = 2) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * es};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, es == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return false; }
    return true;
}

void choose_tile(int B, int H, int W, int* TW, int* TH, int* TB, bool pool = false) {
    static const int cand[][3] = {{16, 8, 1}, {8, 16, 1}, {32, 4, 1}, {4, 32, 1}, {64, 2, 1}, {128, 1, 1}, {8, 8, 2},
                                  {16, 4, 2}, {4, 16, 2}, {8, 4, 4}, {4, 8, 4}, {4, 4, 8}, {8, 2, 8}, {2, 2, 32}, {1, 1, 128}};
    int64_t best = -1;
    for (auto& c : cand) {
        if (pool && !((c[0] % 2) == 0 && c[0] <= 16 && (c[1] % 2) == 0)) continue;   // 2x2 windows must stay inside one warp
        const int64_t tiles = (int64_t)ceil_div(W, c[0]) * ceil_div(H, c[1]) * ceil_div(B, c[2]);
        if (best < 0 || tiles < best) { best = tiles; *TW = c[0]; *TH = c[1]; *TB = c[2]; }
    }
}

template <int BN, int PASSES, bool FP16>
int launch_inst(const TcConvPlan* pl, cudaStream_t s);

}  // namespace

struct TcConvPlan {
    bool two_cta = false;
    bool c64 = false;      // 64 -> 64 channel 3x3 specialisation (conv_c64_kernel)
    bool c64x2 = false;    // ... on a CTA pair (conv_c64x2_kernel)
    bool c1f = false;      // ... with conv1_1 fused in (conv_c1f_kernel; implies c64x2 geometry)
    int device = 0;

    TcConvDesc d;
    CUtensorMap map_x_hi, map_x_lo, map_w_hi, map_w_lo, map_x_h8, map_w_l8;
    CUtensorMap map_y_hi, map_y_lo;      // conv_c64x2_kernel with tma_out
    TcParams p;
    int BN, grid;
    int* err_flag;
};

// SM count and the shared-memory opt-in are per DEVICE (one process may hold contexts on several GPUs)
constexpr int kMaxDevices = 64;
static int current_device() { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; } return dev; }
int tc_num_sms() {
    static int n[kMaxDevices] = {};
    const int dev = current_device() % kMaxDevices;
    if (!n[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 148; }
        n[dev] = v;
    }
    return n[dev];
}
// once per (kernel instance, device): raise the dynamic shared-memory limit
template <typename K>
static int smem_opt_in(K kernel, int bytes, bool* done /*[kMaxDevices]*/) {
    const int dev = current_device() % kMaxDevices;
    if (!done[dev]) {
        H3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done[dev] = true;
    }
    return H3D_OK;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait above); tune.pdl = 0 gives plain stream order.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tc_tuning().pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Tuning switches (A/B experiments, forced kernel variants in the tests).  Read from the environment ONCE, when the library
// is first used, and changeable afterwards only through tc_set_tuning() (h3d_set_tuning): nothing on a launch path calls getenv.
TcTuning& tc_tuning() {
    static TcTuning t = [] {
        TcTuning v;
        auto geti = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
        v.two_cta = geti("H3D_TC_2CTA", -1);
        v.bn = geti("H3D_TC_BN", 0);
        v.c64 = geti("H3D_TC_C64", 1);
        v.c64x2 = geti("H3D_TC_C64X2", 1);
        v.pair128 = geti("H3D_TC_PAIR128", 1);
        v.stack = geti("H3D_TC_STACK", 1);
        v.chunk_kb = geti("H3D_TC_CHUNK_KB", 0);
        v.exp = geti("H3D_TC_EXP", 0);
        v.no_side_stream = geti("H3D_NO_SIDE_STREAM", 0);
        v.no_pool_fusion = geti("H3D_NO_POOL_FUSION", 0);
        v.lift_direct = geti("H3D_LIFT_DIRECT", 0);
        v.c3_ffma = geti("H3D_C3_FFMA", 0);
        v.c3_tma = geti("H3D_C3_TMA", 1);
        v.pdl = geti("H3D_PDL", 1);
        v.fc_chain = geti("H3D_FC_CHAIN", 1);
        v.c64_tma_out = geti("H3D_C64_TMA_OUT", 1);
        v.chain = geti("H3D_TC_CHAIN", 1);
        v.small_batch_split = geti("H3D_TC_SMALL_SPLIT", 1);
        v.fuse_c1 = geti("H3D_FUSE_C1", 1);
        v.no_seg_fusion = geti("H3D_NO_SEG_FUSION", 0);
        return v;
    }();
    return t;
}
int tc_set_tuning(const char* key, int value) {
    TcTuning& t = tc_tuning();
    const std::string k(key ? key : "");
    if (k == "tc_2cta") t.two_cta = value;
    else if (k == "tc_bn") t.bn = value;
    else if (k == "tc_c64") t.c64 = value;
    else if (k == "tc_c64x2") t.c64x2 = value;
    else if (k == "tc_pair128") t.pair128 = value;
    else if (k == "tc_stack") t.stack = value;
    else if (k == "tc_chunk_kb") t.chunk_kb = value;
    else if (k == "tc_exp") t.exp = value;
    else if (k == "no_side_stream") t.no_side_stream = value;
    else if (k == "no_pool_fusion") t.no_pool_fusion = value;
    else if (k == "lift_direct") t.lift_direct = value;
    else if (k == "c3_ffma") t.c3_ffma = value;
    else if (k == "c3_tma") t.c3_tma = value;
    else if (k == "pdl") t.pdl = value;
    else if (k == "fc_chain") t.fc_chain = value;
    else if (k == "c64_tma_out") t.c64_tma_out = value;
    else if (k == "tc_chain") t.chain = value;
    else if (k == "tc_small_split") t.small_batch_split = value;
    else if (k == "fuse_c1") t.fuse_c1 = value;
    else if (k == "no_seg_fusion") t.no_seg_fusion = value;
    else { set_error("h3d_set_tuning: unknown key '%s'", k.c_str()); return H3D_EINVAL; }
    return H3D_OK;
}

namespace {
template <int BN, int PASSES, bool FP16>
int launch_inst(const TcConvPlan* pl, cudaStream_t s) {
    constexpr int smem = num_stages(BN, PASSES) * stage_bytes(BN, PASSES) + 1024 /*align slack*/ + 256 /*barriers*/;
    static bool attr[kMaxDevices] = {};
    if (int rc = smem_opt_in(conv_tc_kernel<BN, PASSES, FP16>, smem, attr)) return rc;
    H3D_CUDA(launch_pdl(conv_tc_kernel<BN, PASSES, FP16>, dim3(pl->grid), dim3(kThreads), smem, s, pl->map_x_hi, pl->map_x_lo, pl->map_w_hi,
                        pl->map_w_lo, pl->map_x_h8, pl->map_w_l8, pl->p));
    return H3D_OK;
}
}  // namespace

namespace {
template <int PASSES, bool FP16>
int launch_c64(const TcConvPlan* pl, cudaStream_t s) {
    constexpr int smem = c64_smem_bytes(PASSES);
    static bool attr[kMaxDevices] = {};
    if (int rc = smem_opt_in(conv_c64_kernel<PASSES, FP16>, smem, attr)) return rc;
    H3D_CUDA(launch_pdl(conv_c64_kernel<PASSES, FP16>, dim3(pl->grid), dim3(kThreads), smem, s, pl->map_x_hi, pl->map_x_lo, pl->map_w_hi, pl->map_w_lo,
                        pl->p));
    return H3D_OK;
}
template <bool FP16>
int launch_c1f(const TcConvPlan* pl, const float* image, cudaStream_t s) {
    static bool attr[kMaxDevices] = {};
    if (int rc = smem_opt_in(conv_c1f_kernel<FP16>, C1F_SMEM, attr)) return rc;
    H3D_CUDA(launch_pdl(conv_c1f_kernel<FP16>, dim3(pl->grid), dim3(C1F_THREADS), (size_t)C1F_SMEM, s, image, pl->map_w_hi, pl->map_w_lo, pl->p));
    return H3D_OK;
}
template <bool FP16>
int launch_c64x2(const TcConvPlan* pl, cudaStream_t s) {
    static bool attr[kMaxDevices] = {};
    if (int rc = smem_opt_in(conv_c64x2_kernel<FP16>, C64X2_SMEM, attr)) return rc;
    H3D_CUDA(launch_pdl(conv_c64x2_kernel<FP16>, dim3(pl->grid), dim3(kThreads2), (size_t)C64X2_SMEM, s, pl->map_x_hi, pl->map_x_lo, pl->map_w_hi,
                        pl->map_w_lo, pl->map_y_hi, pl->map_y_lo, pl->p));
    return H3D_OK;
}
}  // namespace

namespace {
template <int BN, int PASSES, bool FP16>
int launch_inst2(const TcConvPlan* pl, cudaStream_t s) {
    constexpr int smem = num_stages2(BN, PASSES) * stage_bytes2(BN, PASSES) + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(smem <= 227 * 1024, "conv_tc2_kernel: shared memory budget");
    static bool attr[kMaxDevices] = {};
    if (int rc = smem_opt_in(conv_tc2_kernel<BN, PASSES, FP16>, smem, attr)) return rc;
    H3D_CUDA(launch_pdl(conv_tc2_kernel<BN, PASSES, FP16>, dim3(pl->grid), dim3(kThreads2), smem, s, pl->map_x_hi, pl->map_x_lo, pl->map_w_hi,
                        pl->map_w_lo, pl->map_x_h8, pl->map_w_l8, pl->p));
    return H3D_OK;
}
}  // namespace

TcConvPlan* tc_conv_plan_create(const TcConvDesc& d) {
    if (d.Cin_pad % BK != 0 || d.Cout_pad % 64 != 0 || (d.k != 1 && d.k != 3 && d.k != 5 && d.k != 7) || (d.passes != 1 && d.passes != 3 && d.passes != 4)) {
        set_error("tc_conv: unsupported geometry (Cin_pad=%d Cout_pad=%d k=%d passes=%d)", d.Cin_pad, d.Cout_pad, d.k, d.passes);
        return nullptr;
    }
    const bool padded_out = d.Cout % 32 != 0;   // masked scalar tail in the epilogue: no alignment requirement there
    if (d.y.hi && ((d.Cy_total % 8) || (d.cy_off % 8))) { set_error("tc_conv: split output channel offset/stride must be multiples of 8"); return nullptr; }
    if (d.yf && !padded_out && ((d.Cyf_total % 4) || (d.cyf_off % 4))) { set_error("tc_conv: fp32 output channel offset/stride must be multiples of 4"); return nullptr; }
    if (padded_out && d.pool == 1) { set_error("tc_conv: fused pooling needs Cout %% 32 == 0"); return nullptr; }
    if (d.pool < 0 || d.pool > 2) { set_error("tc_conv: pool mode must be 0 (none), 1 (max-pool) or 2 (stride 2)"); return nullptr; }
    if (d.pool == 2 && d.k < 3) { set_error("tc_conv: stride 2 needs k >= 3 (for k = 1 TF's 'SAME' samples the even pixels, not the odd ones)"); return nullptr; }
    if (d.passes == 3 && ((!d.x.lo && !d.c1_w) || !d.w.lo)) { set_error("tc_conv: 3-pass mode needs lo planes"); return nullptr; }
    if (d.passes == 4 && (!d.x.l8 || !d.x.h8 || !d.w.l8 || !d.w.h8 || d.half != Half16::FP16 || d.corr_scale <= 0.f)) {
        set_error("tc_conv: fp8-correction mode needs fp16 + e4m3 l8/h8 planes for activations and weights and a correction scale");
        return nullptr;
    }
    if (d.passes == 4 && d.y.hi && (!d.y.l8 || !d.y.h8)) { set_error("tc_conv: fp8-correction mode needs l8/h8 output planes"); return nullptr; }
    if (d.passes == 4 && d.y.hi && ((d.Cy_total % 16) || (d.cy_off % 16))) { set_error("tc_conv: fp8 planes need 16-channel aligned offsets"); return nullptr; }
    TcConvPlan* pl = new TcConvPlan();
    pl->d = d;
    const TcTuning& tune = tc_tuning();
    int BN = d.Cout_pad % 128 == 0 ? 128 : 64;
    // CTA-pair kernel (cta_group::2, UMMA 256 x BN) when N can be 256, and - in the 3-pass modes - its N-stacked variant for
    // Cout = 128 (conv2_2, the 7x7 refinement layers, conv4_7 / conv5_2): half the weight rows and a quarter fewer operand bytes
    // per SM and K step than the single-CTA stacked kernel, which is bound by shared-memory bandwidth there.  tune.two_cta = 0 / 1
    // forces one kernel family.
    bool two = d.Cout_pad % 256 == 0 || (tune.pair128 && d.passes == 3 && d.Cout_pad % 128 == 0);
    if (tune.two_cta >= 0) two = tune.two_cta != 0;
    if (two) BN = d.Cout_pad % 256 == 0 ? 256 : (d.Cout_pad % 128 == 0 ? 128 : 64);   // CTA pair: UMMA 256 x BN
    // Small maps (lifting pyramids from 16x16 down, the FC stacks = 1x1 convolutions over batch rows): too few pixel tiles to fill
    // the machine with wide tiles, so N = 64 tiles on single CTAs spread the work (and the weight stream) over 4-8x more SMs.  The
    // rule depends on the layer geometry only, never on the batch size: the arithmetic of an image must not depend on how a batch
    // is cut (tests/test_gpu_properties.py: bit-identical results under sharding).
    if (tune.two_cta < 0 && (int64_t)d.H * d.W <= 256) { two = false; BN = 64; }
    if ((tune.bn == 64 || tune.bn == 128 || tune.bn == 256) && d.Cout_pad % tune.bn == 0) BN = tune.bn;
    // 64 -> 64 / 64 -> 128 channels, 3x3 (conv1_2, conv2_1, small lifting layers): weights-resident / patch-reuse kernels; on a CTA
    // pair in the 3-pass modes when the map is large enough to fill the machine with tile pairs
    bool c64 = d.k == 3 && d.Cin_pad == 64 && d.Cout_pad <= 128 && (d.passes == 1 || d.passes == 3) && tune.c64 != 0;
    const bool c64x2 = c64 && d.passes == 3 && tune.c64x2 != 0 && (int64_t)d.H * d.W > 256;
    const bool c1f = d.c1_w != nullptr;
    if (c1f && !(c64x2 && tc_conv_can_fuse_first(d.H, d.W, d.Cin_pad, d.Cout_pad, d.k, d.passes, d.pool) && d.c1_bias && d.y.hi && d.y.lo && !d.yf)) {
        set_error("tc_conv: conv1_1 can only be fused into a pooled 64 -> 64 channel 3x3 layer of a 3-pass mode on a map larger than 16x16");
        delete pl; return nullptr;
    }
    if (c64) { two = false; BN = 64; }
    int TW, TH, TB;
    if (d.pool && ((d.H | d.W) & 1)) { set_error("tc_conv: fused max-pool / stride 2 needs even H and W"); delete pl; return nullptr; }
    choose_tile(d.B, d.H, d.W, &TW, &TH, &TB, d.pool == 1);
    if (c64) { TW = C64_TW; TH = C64_TH; TB = 1; }
    // Few pixel tiles (small batches: run.py's single image, BASELINE config 1): a layer is then as slow as ONE of its work items, and the
    // CTA pair's 256 x 256 items are the longest there are (72 K blocks x 1536 tensor cycles = 60 us for conv4_x) while most SMs idle.
    // When the pair kernel would occupy at most half of the machine, the same layer runs as N = 64 (else N = 128) tiles on single CTAs -
    // 4x (2x) shorter items on 4x (2x) more SMs.  The instruction ORDER per output element is kept (N-stacked iff the pair kernel would
    // have been N-stacked, same K chunks), so the result is bit-identical to the pair kernel's and still independent of the batch size
    // (tests/test_gpu_properties.py: a batch of 32 equals its shards of 1 + 1 + 30 bit for bit).
    bool stack_single = tune.stack != 0;
    if (two && !c64 && tune.two_cta < 0 && tune.bn == 0 && tune.small_batch_split) {
        const int tiles = ceil_div(d.W, TW) * ceil_div(d.H, TH) * ceil_div(d.B, TB);
        const int sms = tc_num_sms();
        const int pair_items = ceil_div(tiles, 2) * (d.Cout_pad / BN);
        if (4 * pair_items <= sms) {
            const bool pair_stacked = stack2(BN, d.passes);
            int bn1 = 0;
            if (tiles * (d.Cout_pad / 64) <= sms) bn1 = 64;
            else if (d.Cout_pad % 128 == 0 && tiles * (d.Cout_pad / 128) <= sms) bn1 = 128;
            if (bn1) { two = false; BN = bn1; stack_single = pair_stacked; }
        }
    }
    pl->BN = BN;
    pl->two_cta = two;
    pl->c64 = c64;
    pl->c64x2 = c64x2;
    pl->c1f = c1f;
    pl->device = current_device();
    TcParams& p = pl->p;
    p.bias = d.bias;
    p.y_hi = d.y.hi; p.y_lo = d.y.lo; p.y_l8 = d.y.l8; p.y_h8 = d.y.h8; p.Cy_total = d.Cy_total; p.cy_off = d.cy_off;
    p.corr_scale = d.corr_scale;
    p.yf = d.yf; p.Cyf_total = d.Cyf_total; p.cyf_off = d.cyf_off;
    p.B = d.B; p.H = d.H; p.W = d.W; p.k = d.k; p.pad = d.k / 2; p.cin_chunks = d.Cin_pad / BK;
    p.TW = TW; p.TH = TH; p.TB = TB;
    p.tiles_w = ceil_div(d.W, TW); p.tiles_h = ceil_div(d.H, TH);
    const int tiles_b = ceil_div(d.B, TB);
    p.n_tiles = d.Cout_pad / BN;
    p.num_tiles = p.tiles_w * p.tiles_h * tiles_b * p.n_tiles;
    if (two) p.num_tiles = ceil_div(p.tiles_w * p.tiles_h * tiles_b, 2) * p.n_tiles;   // work items = pixel-tile PAIRS x N tiles
    p.leaky = d.leaky;
    p.n_valid = d.Cout;
    p.pool = d.pool;
    p.stack = stack_single;
    p.exp = tune.exp;
    p.err_flag = d.err_flag;
    p.w1 = d.c1_w; p.bias1 = d.c1_bias; p.leaky1 = d.c1_leaky;
    // <= ~108 accumulating MMAs per TMEM partial sum (9 K blocks x 4 K steps x 3 passes); BN = 256 keeps everything in
    // TMEM (its 256 fp32 partial sums per thread would not fit the register file)
    p.chunk_kb = d.passes >= 3 ? 9 : 27;
    if (tune.chunk_kb > 0) p.chunk_kb = tune.chunk_kb;
    if (BN > 128 && !two) p.chunk_kb = 1 << 30;
    pl->grid = two ? 2 * std::min(p.num_tiles, tc_num_sms() / 2) : std::min(p.num_tiles, tc_num_sms());
    if (c64) {   // work items = pixel tiles; CTAs are split evenly over the 64-channel groups
        p.num_tiles = p.tiles_w * p.tiles_h * tiles_b;
        pl->grid = p.n_tiles * std::max(1, std::min(p.num_tiles, tc_num_sms() / p.n_tiles));
    }
    if (c64x2) {   // work items = pixel-tile pairs; clusters are split evenly over the 64-channel groups
        p.num_tiles = ceil_div(p.tiles_w * p.tiles_h * tiles_b, 2);
        pl->grid = 2 * p.n_tiles * std::max(1, std::min(p.num_tiles, (tc_num_sms() / 2) / p.n_tiles));
    }
    const bool stacked_pair = two && stack2(BN, d.passes);
    const int w_box_rows = c64x2 ? 32 : stacked_pair ? BN : two ? BN / 2 : BN;
    const int Ktot = d.k * d.k * d.Cin_pad;
    const int box_h = c64 ? C64_PH : TH;   // the 64 -> 64 kernel fetches the tile rows plus the halo rows in one box
    bool ok;
    if (c1f) {   // the activation patches are computed in the kernel: weight maps only
        ok = encode_w_map(&pl->map_w_hi, d.w.hi, Ktot, d.Cout_pad, w_box_rows) && encode_w_map(&pl->map_w_lo, d.w.lo, Ktot, d.Cout_pad, w_box_rows);
        pl->map_x_hi = pl->map_w_hi; pl->map_x_lo = pl->map_w_lo;
    } else {
        ok = encode_act_map(&pl->map_x_hi, d.x.hi, d.Cin_total, d.Cin_pad, d.W, d.H, d.B, TW, box_h, TB) &&
             encode_w_map(&pl->map_w_hi, d.w.hi, Ktot, d.Cout_pad, w_box_rows);
        if (ok && d.passes == 3)
            ok = encode_act_map(&pl->map_x_lo, d.x.lo, d.Cin_total, d.Cin_pad, d.W, d.H, d.B, TW, box_h, TB) &&
                 encode_w_map(&pl->map_w_lo, d.w.lo, Ktot, d.Cout_pad, w_box_rows);
    }
    if (ok && d.passes == 4)   // e4m3 planes: x residual (slot "lo"), x coarse, w coarse (slot "lo"), w residual
        ok = encode_act_map(&pl->map_x_lo, d.x.l8, d.Cin_total, d.Cin_pad, d.W, d.H, d.B, TW, TH, TB, 1) &&
             encode_act_map(&pl->map_x_h8, d.x.h8, d.Cin_total, d.Cin_pad, d.W, d.H, d.B, TW, TH, TB, 1) &&
             encode_w_map(&pl->map_w_lo, d.w.h8, Ktot, d.Cout_pad, w_box_rows, 1) &&
             encode_w_map(&pl->map_w_l8, d.w.l8, Ktot, d.Cout_pad, w_box_rows, 1);
    if (ok && d.passes == 1) { pl->map_x_lo = pl->map_x_hi; pl->map_w_lo = pl->map_w_hi; }
    if (ok && d.passes != 4) { pl->map_x_h8 = pl->map_x_hi; pl->map_w_l8 = pl->map_w_hi; }
    if (ok && stacked_pair) ok = encode_w_map(&pl->map_w_l8, d.w.hi, Ktot, d.Cout_pad, BN / 2);   // region X: W_hi in 64-row boxes
    p.tma_out = 0;
    pl->map_y_hi = pl->map_x_hi; pl->map_y_lo = pl->map_x_hi;
    if (ok && c64x2 && d.pool == 0 && d.y.hi && d.y.lo && !d.yf && d.Cout % 64 == 0 && tune.c64_tma_out) {
        // output planes [B,H,W,Cy_total] at channel offset cy_off: boxes of {64 channels, 16, 8, 1} at channel coordinate n0
        ok = encode_act_map(&pl->map_y_hi, d.y.hi + d.cy_off, d.Cy_total, d.Cout_pad, d.W, d.H, d.B, C64_TW, C64_TH, 1) &&
             encode_act_map(&pl->map_y_lo, d.y.lo + d.cy_off, d.Cy_total, d.Cout_pad, d.W, d.H, d.B, C64_TW, C64_TH, 1);
        p.tma_out = 1;
    }
    if (!ok) { delete pl; return nullptr; }
    return pl;
}

void tc_conv_plan_destroy(TcConvPlan* p) { delete p; }

// Layer chains (conv_tc2_kernel): only the CTA-pair kernel draws tickets / signals / waits per image
bool tc_conv_plan_chainable(const TcConvPlan* p) { return p->two_cta && !p->c64 && !p->c64x2; }
int tc_conv_plan_signal_target(const TcConvPlan* p) { return p->p.tiles_w * p->p.tiles_h * p->p.n_tiles; }
const TcConvDesc& tc_conv_plan_desc(const TcConvPlan* p) { return p->d; }
void tc_conv_plan_set_chain(TcConvPlan* p, int* sched, const int* dep_cnt, int dep_target, int* sig_cnt) {
    p->p.sched = sched; p->p.dep_cnt = dep_cnt; p->p.dep_target = dep_target; p->p.sig_cnt = sig_cnt;
}

int64_t tc_conv_flops(const TcConvPlan* p) {
    return 2ll * p->d.B * p->d.H * p->d.W * p->d.k * p->d.k * (int64_t)p->d.Cin_pad * p->d.Cout_pad;
}

bool tc_conv_can_fuse_first(int H, int W, int Cin, int Cout, int k, int passes, int pool) {
    const TcTuning& t = tc_tuning();
    return t.fuse_c1 && t.c64 && t.c64x2 && t.two_cta < 0 && t.bn == 0 && k == 3 && Cin == 64 && Cout == 64 && passes == 3 && pool == 1 &&
           (int64_t)H * W > 256 && (H % 2) == 0 && (W % 2) == 0;
}

int tc_conv_launch_image(const TcConvPlan* pl, const float* image, cudaStream_t s) {
    H3D_REQUIRE(pl->c1f && image, "tc_conv_launch_image: the plan has no fused first layer");
    return pl->d.half == Half16::FP16 ? launch_c1f<true>(pl, image, s) : launch_c1f<false>(pl, image, s);
}

int tc_conv_launch(const TcConvPlan* pl, cudaStream_t s) {
    const bool fp16 = pl->d.half == Half16::FP16;
    H3D_REQUIRE(!pl->c1f, "tc_conv_launch: the plan fuses the first layer, launch it with the image (tc_conv_launch_image)");
    const int key = pl->BN * 10 + pl->d.passes;
    if (pl->c64x2) return fp16 ? launch_c64x2<true>(pl, s) : launch_c64x2<false>(pl, s);
    if (pl->c64) {
        if (pl->d.passes == 3) return fp16 ? launch_c64<3, true>(pl, s) : launch_c64<3, false>(pl, s);
        return fp16 ? launch_c64<1, true>(pl, s) : launch_c64<1, false>(pl, s);
    }
    if (pl->two_cta) {
#define CASE2(BN_, P_)                                                                 \
    case BN_ * 10 + P_:                                                                \
        return fp16 ? launch_inst2<BN_, P_, true>(pl, s) : launch_inst2<BN_, P_, false>(pl, s);
        switch (key) {
            CASE2(64, 1) CASE2(64, 3) CASE2(128, 1) CASE2(128, 3) CASE2(256, 1) CASE2(256, 3)
            case 64 * 10 + 4: return launch_inst2<64, 4, true>(pl, s);
            case 128 * 10 + 4: return launch_inst2<128, 4, true>(pl, s);
            case 256 * 10 + 4: return launch_inst2<256, 4, true>(pl, s);
        }
#undef CASE2
    }
#define CASE(BN_, P_)                                                                  \
    case BN_ * 10 + P_:                                                                \
        return fp16 ? launch_inst<BN_, P_, true>(pl, s) : launch_inst<BN_, P_, false>(pl, s);
    switch (key) {
        CASE(64, 1) CASE(64, 3) CASE(128, 1) CASE(128, 3) CASE(256, 1) CASE(256, 3)
        case 64 * 10 + 4: return launch_inst<64, 4, true>(pl, s);      // fp16 + e4m3 corrections
        case 128 * 10 + 4: return launch_inst<128, 4, true>(pl, s);
        case 256 * 10 + 4: return launch_inst<256, 4, true>(pl, s);
    }
#undef CASE
    set_error("tc_conv: no kernel instance for BN=%d passes=%d", pl->BN, pl->d.passes);
    return H3D_EINVAL;
}

// conv1_1 (3 -> 64 channels, 3x3, stride 1) on the tensor cores: x fp32 [B,H,W,3], w fp32 HWIO [3,3,3,64] and bias [64] on the
// device, output split planes y (hi, and lo when present) [B,H,W,Cs_total] at channel offset cs_off.
int launch_conv_c3_tc(const float* x, const float* w, const float* bias, Split y, int Cs_total, int cs_off, int B, int H, int W, int leaky,
                      Half16 half, cudaStream_t s, int* err_flag) {
    H3D_REQUIRE(x && w && bias && y.hi && !y.l8 && (Cs_total % 8) == 0 && (cs_off % 8) == 0, "conv_c3_tc: bad argument");
    TcParams p{};
    p.bias = bias;
    p.y_hi = y.hi; p.y_lo = y.lo; p.Cy_total = Cs_total; p.cy_off = cs_off;
    p.corr_scale = 1.f;
    p.B = B; p.H = H; p.W = W; p.k = 3; p.pad = 1; p.cin_chunks = 1;
    p.TW = C64_TW; p.TH = C64_TH; p.TB = 1;
    p.tiles_w = ceil_div(W, C64_TW); p.tiles_h = ceil_div(H, C64_TH); p.n_tiles = 1;
    p.num_tiles = p.tiles_w * p.tiles_h * B;
    p.n_valid = 64; p.pool = 0; p.chunk_kb = 1; p.leaky = leaky; p.err_flag = err_flag;
    const int grid = std::min(p.num_tiles, 2 * tc_num_sms());   // two co-resident CTAs per SM (86 KB, 72 registers, 256 TMEM columns each)
    if (tc_tuning().c3_tma) {   // smem-staged epilogue with bulk tensor stores (output tensor maps over the 64-channel slice of the planes)
        CUtensorMap my_hi, my_lo;
        if (!encode_act_map(&my_hi, y.hi + cs_off, Cs_total, 64, W, H, B, C64_TW, C64_TH, 1)) return H3D_ECUDA;
        my_lo = my_hi;
        if (y.lo && !encode_act_map(&my_lo, y.lo + cs_off, Cs_total, 64, W, H, B, C64_TW, C64_TH, 1)) return H3D_ECUDA;
        static bool at_h[kMaxDevices] = {}, at_b[kMaxDevices] = {};
        if (int rc = smem_opt_in(conv_c3_tma_kernel<true>, C3S_SMEM, at_h)) return rc;
        if (int rc = smem_opt_in(conv_c3_tma_kernel<false>, C3S_SMEM, at_b)) return rc;
        if (half == Half16::FP16) H3D_CUDA(launch_pdl(conv_c3_tma_kernel<true>, dim3(grid), dim3(C3T_THREADS), (size_t)C3S_SMEM, s, x, w, my_hi, my_lo, p));
        else H3D_CUDA(launch_pdl(conv_c3_tma_kernel<false>, dim3(grid), dim3(C3T_THREADS), (size_t)C3S_SMEM, s, x, w, my_hi, my_lo, p));
        return H3D_OK;
    }
    static bool attr_h[kMaxDevices] = {}, attr_b[kMaxDevices] = {};
    if (int rc = smem_opt_in(conv_c3_tc_kernel<true>, C3T_SMEM, attr_h)) return rc;
    if (int rc = smem_opt_in(conv_c3_tc_kernel<false>, C3T_SMEM, attr_b)) return rc;
    if (half == Half16::FP16) H3D_CUDA(launch_pdl(conv_c3_tc_kernel<true>, dim3(grid), dim3(C3T_THREADS), (size_t)C3T_SMEM, s, x, w, p));
    else H3D_CUDA(launch_pdl(conv_c3_tc_kernel<false>, dim3(grid), dim3(C3T_THREADS), (size_t)C3T_SMEM, s, x, w, p));
    return H3D_OK;
}

// ------------------------------------------------------------------------------------------ FC chain: host
struct FcChainPlan { FcChainParams P; Half16 half; int grid; };

FcChainPlan* fc_chain_plan_create(const FcChainDesc* chains, int num_chains, int B, Half16 half, const float* can, const float* uxyz,
                                  unsigned int* counter, int* err_flag) {
    if (num_chains < 1 || num_chains > 2 || B < 1) { set_error("fc_chain: bad geometry"); return nullptr; }
    FcChainPlan* pl = new FcChainPlan();
    memset(&pl->P, 0, sizeof(pl->P));
    pl->half = half; pl->grid = num_chains * kFcCluster;
    FcChainParams& P = pl->P;
    P.num_chains = num_chains; P.B = B; P.can = can; P.uxyz = uxyz; P.counter = counter; P.err_flag = err_flag;
    for (int c = 0; c < num_chains; ++c) {
        const FcChainDesc& cd = chains[c];
        if (cd.num_layers < 1 || cd.num_layers > kFcMaxLayers) { set_error("fc_chain: 1..4 layers per chain"); delete pl; return nullptr; }
        P.chain[c].num_layers = cd.num_layers;
        for (int l = 0; l < cd.num_layers; ++l) {
            const FcLayerDesc& d = cd.layer[l];
            FcLayer& L = P.chain[c].layer[l];
            const int Kpad = (int)align_up(d.in_features, BK);
            if (!d.x.hi || !d.x.lo || !d.w.hi || !d.w.lo || d.out_pad % 64 || d.x_stride < Kpad || (d.y.hi && (d.y_stride % 8))) {
                set_error("fc_chain: bad layer %d of chain %d", l, c); delete pl; return nullptr;
            }
            bool ok = encode_act_map(&L.map_x_hi, d.x.hi, d.x_stride, Kpad, 1, 1, B, 1, 1, BM) &&
                      encode_act_map(&L.map_x_lo, d.x.lo, d.x_stride, Kpad, 1, 1, B, 1, 1, BM) &&
                      encode_w_map(&L.map_w_hi, d.w.hi, Kpad, d.out_pad, 64) && encode_w_map(&L.map_w_lo, d.w.lo, Kpad, d.out_pad, 64);
            if (!ok) { delete pl; return nullptr; }
            L.kblocks = Kpad / BK; L.m_tiles = ceil_div(B, BM); L.n_tiles = d.out_pad / 64;
            TcParams& p = L.p;
            p.bias = d.bias; p.y_hi = d.y.hi; p.y_lo = d.y.lo; p.Cy_total = d.y_stride; p.cy_off = 0; p.corr_scale = 1.f;
            p.yf = d.yf; p.Cyf_total = d.yf_stride; p.cyf_off = 0;
            p.B = B; p.H = 1; p.W = 1; p.k = 1; p.TW = 1; p.TH = 1; p.TB = BM;
            p.n_valid = d.out_features; p.pool = 0; p.leaky = d.leaky; p.stack = 1; p.err_flag = err_flag;
        }
    }
    return pl;
}

void fc_chain_plan_destroy(FcChainPlan* p) { delete p; }

int fc_chain_launch(const FcChainPlan* pl, const float* hand_side, float* rot, float* out, cudaStream_t s) {
    FcChainParams P = pl->P;
    P.hand_side = hand_side; P.rot = rot; P.out = out;
    H3D_REQUIRE(P.num_chains == 1 || (hand_side && out), "fc_chain: hand_side / out are required for the rotation epilogue");
    constexpr int smem = num_stages(64, 3) * stage_bytes(64, 3) + 1024 + 256;
    static bool at_h[kMaxDevices] = {}, at_b[kMaxDevices] = {};
    if (pl->half == Half16::FP16) {
        if (int rc = smem_opt_in(fc_chain_kernel<true>, smem, at_h)) return rc;
        H3D_CUDA(launch_pdl(fc_chain_kernel<true>, dim3(pl->grid), dim3(kThreads), (size_t)smem, s, P));
    } else {
        if (int rc = smem_opt_in(fc_chain_kernel<false>, smem, at_b)) return rc;
        H3D_CUDA(launch_pdl(fc_chain_kernel<false>, dim3(pl->grid), dim3(kThreads), (size_t)smem, s, P));
    }
    return H3D_OK;
}

}  // namespace h3d

// ALS-NMF hot loop on B200 (sm_100a): replaces als_nmf's `while (i < maxIter)` body
// (R/bicluster.R:119-203) and the two NMF::.fcnnls calls inside it (:123, :150).
//
// One ALS iteration = two half-steps, each a STREAM launch followed by a SOLVE launch:
//   H-step:  R = W'A  (k x n), then exact NNLS per column of A   -> H
//   W-step:  R = A H' (m x k), then exact NNLS per row of A      -> W,
//            plus residNorm / rms(dW) / rms(dH) and the convergence test.
// Stream kernel (als_stream_kernel): A (m x n, column-major, FP64) is read from HBM exactly once per half-step:
//   * a producer warp moves 16-row x 128-col (H-step) or 128-row x 32-col (W-step) boxes of A, and the matching
//     slab of the fixed factor, into a shared-memory ring with TMA (cp.async.bulk.tensor, SWIZZLE_128B)
//     signalled on mbarriers;
//   * two groups of four consumer warps, one stage out of phase, read the m8n8k4 FP64 tensor-core fragments
//     straight out of the swizzled ring (conflict-free by construction) and accumulate with DMMA.8x8x4;
//   * work is cut stream-K style: the (tile, reduction-stage) space is split into 148 equal contiguous
//     ranges, one persistent CTA per SM; a consumer group writes one k x 128 partial per tile it touches.
// Solve kernel (als_solve_kernel): thousands of warps sum the partials in slot order (deterministic) and run
// the exact NNLS for 16 units per CTA, write the factor, and fold the Gram of the new factor, its inverse and
// the iteration scalars in two fixed-order levels.  Splitting the launches exposes one launch gap per
// half-step but replaces a tail of ~40 (H) / ~100-200 (W) us during which 8 warps per SM ran the NNLS of a whole
// 128-unit tile with a ~15 us kernel at full occupancy (profiles/r01_*).
// Nothing but A's two passes touches HBM in bulk: factors, partials and Grams stay in L2.
//
// Row-sharded sessions (mfb_als_begin_sharded): A and W are split by rows over several devices,
// H is replicated.  The same kernels run on every rank's block; W_r'A_r, W_r'W_r and three scalars are summed over
// the ranks either by a host-supplied collective between the kernels (callback exchange: als_shard_fold / invert /
// finish kernels) or over cudaIpc-mapped peer memory inside the kernels that consume the sums (ShardPeers: the Gram
// side job, the fold kernel's last block + the H-step solve kernel, the last warp of the W-step solve kernel).
#include <stdlib.h>

#include "common.cuh"
#include "nnls.cuh"

#ifndef ALS_NG
#define ALS_NG 2                  // consumer groups of 4 warps, one stage out of phase with each other
#endif
#define ALS_CONSUMERS (ALS_NG * 128)
#define ALS_THREADS (ALS_CONSUMERS + 128)   // + the producer warpgroup (TMA lane, Gram side job)
#define ALS_STR2(x) #x
#define ALS_STR(x) ALS_STR2(x)
#define TILE_UNITS 128            // columns (H-step) or rows (W-step) per tile

// ---------------------------------------------------------------------------
// device control block (lives in global memory, mirrored to the host after a batch)
// ---------------------------------------------------------------------------
struct AlsCtrl {
    int stop;                  // != 0: every later launch exits immediately
    int reason;                // 0 none, 1 converged, 2 zero row in H, 3 singular W'W, 4 singular HH'
    int iters_done;            // completed (successful) iterations in this batch epoch
    int fixed;                 // 1: no convergence break
    int groups_done[2];        // per half-step completion counters of the solve kernel's CTA groups
    int max_trace;
    int count_draws;           // != 0: the solve kernels count the uniforms .fcnnls' max.col calls would consume
    unsigned long long nzmask; // bit c set when row c of H has a non-zero entry
    double eps;
    double resid_old;
    double normA2;
    double mn, mk, kn;         // element counts for the RMS values
    double dH2;
    double resid, dW, dH;
    unsigned long long draws;  // running total of those draws (ranks up to 32)
    int pending_epoch;         // peer-memory sessions: epoch whose W-step scalars are written but not yet summed (0: none)
    int unpublished_epoch;     // ... and whose flag has not been raised in the peers' blocks yet (0: none)
};

struct HalfParams {
    const double *Fin;  int64_t ldin;     // fixed factor, [KP][ldin]
    long long nfix;                       // valid units of the fixed factor (m for the H-step, n for the W-step)
    double *Fout; const double *Fold; int64_t ldout;   // factor being solved / its previous value
    double *G; double *Minv;              // Gram of Fin and its inverse, KP x KP: written by the stream kernel's side job
    double *gram_part, *gram_part2;       // [stream CTA][KP*KP] / [group][KP*KP] partial Grams of Fin
    int *gram_counter;                    // [1 + groups] arrival counters
    double *P; int maxslots;              // partials [tile][slot][KP][128]
    int *group_counter;                   // [solve groups]
    double *spart, *spart2;               // scalar partials per solve warp / per group, [..][4]
    AlsCtrl *ctrl; double *trace;
    int k, T, SPT;
    long long U;                          // T * SPT
    long long nvalid;                     // number of real units (n for H-step, m for W-step)
    int Gstream;                          // grid size of the stream kernel (stream-K partition)
    int shard;                            // != 0: A is one row block of a matrix sharded over several devices
    double *xscal;                        // shard mode: [4] iteration scalars of the W-step, summed over the ranks
    const struct ShardPeers *peers;       // shard mode over peer memory (NULL: the host's all-reduce callback is used)
    int epoch;                            // shard mode: number of this iteration since the session began (1-based)
    int invert_later;                     // != 0: the stream kernel's side job stops at G; als_shard_invert_kernel inverts it
    double reg_all, reg_diag;             // snmf (R/bicluster.R:341-441): G <- G + reg_all 11' + reg_diag I before it is published / inverted
    int peer_invert_later;                // peer-memory H-step: the side job only publishes W_r'W_r; als_shard_close_invert_kernel sums and inverts
};

// Row-sharded mode over peer memory (NVLink / NVSwitch, cudaIpc mappings): every rank owns one SHARED BLOCK
//   int flagsA[16] | int flagsB[16] | int fold_counter | stamps @192 | double xscal[2][4] @256 | int flagsG[16] @320
//   | double xchg[2][slots][cnt] @512 | double gram[2][64*64]
// mapped into every peer.  Epoch e uses the buffers of parity e & 1.  flagsG[q] = e: rank q's W_q'W_q is in its
// gram buffer; flagsA[q] = e: rank q's folded W_q'A_q is in its exchange buffer; flagsB[q] = e: rank q's W-step
// scalars are complete (all written by rank q into EVERY rank's block with a system-scope release store).  The
// consumers read all ranks' buffers straight over NVLink and add them in rank order, so every rank forms
// bit-identical sums -- a one-shot all-reduce fused into the kernels that need the result, no collective launch:
//   * W'W: the Gram side job of the H-step stream kernel publishes its W_q'W_q early in the launch; the sum over the
//     ranks and the inversion happen AFTER the stream kernel, when every flag has long arrived: in an extra block of
//     the fold kernel (ranks up to 32; next to the folding blocks and their rendezvous, off the critical path) or in
//     als_shard_close_invert_kernel (larger ranks).  (MFB_SHARD_INVERT_IN_STREAM=1: the side job itself waits, sums and
//     inverts while A is streaming, as in round 1 -- on warps that share issue slots with the streaming warps this
//     chain stretched the stream kernel by 10-30 us);
//   * W'A: the fold kernel publishes and waits; the H-step solve kernel sums while it loads its right-hand sides.
//     With at most `slots` ranks the fold kernel PUSHES: it stores its folded W_q'A_q into slot q of every rank's block
//     (posted remote stores, fire and forget), and the solve kernel sums `world` LOCAL buffers instead of pulling
//     (world - 1) x 655 KB over NVLink on its critical path (H-step solve 30 -> 19 us at eight GPUs); with more ranks
//     than slots, slot 0 holds the rank's own buffer and the solve kernel pulls;
//   * scalars: the last warp of the W-step solve kernel publishes; als_shard_close_kernel, placed after the next H-step's
//     stream kernel (and at the end of a batch), waits, sums and closes the iteration.
#define SHARD_MAXP 16
#define SHARD_OFF_FLAGSB 64
#define SHARD_OFF_COUNTER 128
#define SHARD_OFF_STAMPS 192              // 8 x u64 debug time stamps (globaltimer) of the last fold launch
#define SHARD_OFF_XSCAL 256
#define SHARD_OFF_FLAGSG 320
#define SHARD_OFF_XCHG 512
#define SHARD_GRAM_DOUBLES 4096           // 64 x 64
struct ShardPeers {
    int world, rank;
    int slots, push;                      // exchange buffers per parity in every block; push != 0: slot q of every block is written by rank q
    long long cnt;                        // doubles per exchange buffer: T * KP * 128 + KP * KP
    unsigned char *base[SHARD_MAXP];      // base[q] = rank q's shared block as mapped here (base[rank] = local)
};
// exchange buffer `slot` of parity `parity` in rank q's block (pull layout: slot 0 = rank q's own folded buffer)
__device__ __forceinline__ double *shard_xchg(const ShardPeers *sp, int q, int parity, int slot = 0) {
    return (double *)(sp->base[q] + SHARD_OFF_XCHG) + ((size_t)parity * sp->slots + slot) * (size_t)sp->cnt;
}
__device__ __forceinline__ double *shard_gram(const ShardPeers *sp, int q, int parity) {
    return (double *)(sp->base[q] + SHARD_OFF_XCHG) + 2 * (size_t)sp->slots * (size_t)sp->cnt + (size_t)parity * SHARD_GRAM_DOUBLES;
}
__device__ __forceinline__ double *shard_xscal(const ShardPeers *sp, int q, int parity) {
    return (double *)(sp->base[q] + SHARD_OFF_XSCAL) + parity * 4;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(int *p, int v) {
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// publish epoch `e` in flag slot `rank` of the flag array at byte offset `off` of EVERY rank's block: thread q of the
// calling group signals rank q (one fence, then all remote stores in flight at once).  The caller has synchronised
// the group after the data stores (bar.sync / __syncwarp), so they happen-before each signalling thread's fence.
__device__ __forceinline__ void shard_publish(const ShardPeers *sp, int off, int e, int q) {
    if (q < sp->world) {
        __threadfence_system();
        st_relaxed_sys((int *)(sp->base[q] + off) + sp->rank, e);
    }
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must never hang the GPU.  On a time-out the first waiter records who / where
// in a global debug record, raises a global abort flag (every later wait returns at once) and the launch
// drains; the host reports MFB_ERR_CUDA with the record.
__device__ int g_mbar_abort = 0;
__device__ int g_mbar_record[8];
__device__ __noinline__ void mbar_timeout(int code, int a, int b, int c) {
    if (atomicCAS(&g_mbar_abort, 0, 1) == 0) {
        g_mbar_record[0] = code; g_mbar_record[1] = (int)blockIdx.x; g_mbar_record[2] = (int)threadIdx.x;
        g_mbar_record[3] = a; g_mbar_record[4] = b; g_mbar_record[5] = c;
        __threadfence();
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int code = 0, int a = 0, int b = 0) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (*(volatile int *)&g_mbar_abort) return;
        if (clock64() - t0 > 400000000ll) {      // ~0.2 s
            mbar_timeout(code, a, b, (int)parity);
            return;
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int x, int y, int z, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
// keep a loop-invariant value in a register (stops the compiler from re-reading it from the constant bank every iteration)
__device__ __forceinline__ int pin_reg(int v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
// Programmatic dependent launch: let the next kernel of the stream be scheduled as soon as this grid's CTAs have
// started (its CTAs only become resident as ours retire), then wait until the previous grid has completed and its
// writes are visible before touching anything it produced.  Hides the launch latency between the four kernels of
// an iteration; a no-op when the launch does not carry the attribute.
// development aid (mfb_als_timeline): globaltimer of block 0 / thread 0 of the kernels of the last iteration, when the
// CTA was scheduled [2 slot] and when the kernel before it had completed [2 slot + 1]; slots: 0 H stream, 1 close,
// 2 fold, 3 H solve, 4 W stream, 5 W solve; [24], [25] = the H stream of the iteration before
__device__ unsigned long long g_tl[32];
__device__ int g_tl_on = 0;
__device__ __forceinline__ void pdl_prologue(int slot = -1) {
    const bool rec = slot >= 0 && blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && *(volatile int *)&g_tl_on;
    if (rec) {
        if (slot == 0) g_tl[24] = g_tl[0];
        g_tl[2 * slot] = globaltimer_ns();
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (rec) {
        if (slot == 0) g_tl[25] = g_tl[1];
        g_tl[2 * slot + 1] = globaltimer_ns();
    }
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, " ALS_STR(ALS_CONSUMERS) ";" ::: "memory"); }

__device__ __forceinline__ double ldcg(const double *p) { return __ldcg(p); }

// cta index that owns unit u under the balanced partition u0(c) = floor(c * U / G)
__device__ __forceinline__ int cta_of_unit(long long u, long long U, int G) {
    return (int)(((u + 1) * (long long)G + U - 1) / U) - 1;
}

// wait until rank q has published epoch `e` in this rank's flag array (bounded: a dead peer must not hang the GPU;
// on a time-out the launch chain drains with ctrl->stop and the host reports MFB_ERR_CUDA, kind 5)
__device__ __noinline__ void shard_wait_flag(const int *flag, int e, int q, AlsCtrl *ctrl) {
    if (ld_acquire_sys(flag) >= e) return;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < e) {
        if (*(volatile int *)&g_mbar_abort) return;
        if (globaltimer_ns() - t0 > 20000000000ull) {        // 20 s
            mbar_timeout(5, e, q, *(volatile const int *)flag);
            ctrl->stop = 1;
            ctrl->reason = 5;
            return;
        }
        __nanosleep(200);
    }
}

// ---------------------------------------------------------------------------
// the half-step kernel
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// ---------------------------------------------------------------------------
// the solve kernel: second launch of a half-step
// ---------------------------------------------------------------------------
// A CTA owns SU consecutive units (columns of H / rows of W) and, after one barrier, its warps are independent:
// the CTA sums the stream-K partials of its units in slot order (deterministic, coalesced), then every warp solves
// its UPW exact NNLS problems (lane-distributed principal pivoting, nnls.cuh), writes its slice of the factor and
// leaves one partial of the iteration scalars.  Thousands of warps are in flight (10000 at 20000 x 5000, k = 16)
// instead of the 8 per SM a fused tile epilogue could use.  The scalars are folded in two fixed-order levels by the
// last warp of a group of SOLVE_GROUP warps and then by the last group, which also updates the control block.
// (The Gram matrix of the factor produced here is formed by the NEXT stream kernel's side job.)
#ifndef SOLVE_THREADS
#define SOLVE_THREADS 128
#endif
#define SOLVE_WARPS (SOLVE_THREADS / 32)
#ifndef SOLVE_MINB_LARGE
#define SOLVE_MINB_LARGE 3        // resident solve CTAs per SM asked for at ranks 33..56 (register cap 168; k = 64 runs better uncapped)
#endif
#define SOLVE_GROUP 128          // warps per fold group

template <int KB>
struct SolveCfg {
    static constexpr int KP = KB * 8;
    static constexpr int LP = NnlsCfg<KP>::LP;
    static constexpr int R = NnlsCfg<KP>::R;
    static constexpr int UPW = 2;                                // units per warp
    static constexpr int SU = SOLVE_WARPS * UPW;                 // units per CTA
    static constexpr int XS = SU + 1;
    static constexpr size_t SMEM = (size_t)(2 * KP * KP + KP * XS + SOLVE_WARPS * NnlsCfg<KP>::SCR + 8) * sizeof(double);
};

// end of an iteration (R/bicluster.R:173-187): residNorm, rms(dW), rms(dH), trace row, convergence test
__device__ __forceinline__ void finish_iteration(AlsCtrl *c, double *trace, double sd2, double st1, double st2) {
    double r2 = (c->normA2 - 2.0 * st1 + st2) / c->mn;
    const double resid = sqrt(r2 > 0.0 ? r2 : 0.0);
    const double dW = sqrt(sd2 / c->mk);
    const double dH = sqrt(c->dH2 / c->kn);
    const int it = c->iters_done;
    if (it < c->max_trace) {
        trace[it * 3 + 0] = resid;
        trace[it * 3 + 1] = dW;
        trace[it * 3 + 2] = dH;
    }
    c->resid = resid;
    c->dW = dW;
    c->dH = dH;
    c->iters_done = it + 1;
    if (!c->fixed && (fabs(c->resid_old - resid) <= c->eps || (dW <= c->eps && dH <= c->eps))) {
        c->stop = 1;
        c->reason = 1;
    } else {
        c->resid_old = resid;
    }
}

// number of partial slots the stream kernel fills for a tile (see its flush())
__device__ __forceinline__ int expected_slots(const HalfParams &p, int tile) {
    const long long tb = (long long)tile * p.SPT;
    if (p.U >= p.Gstream)
        return ALS_NG * (cta_of_unit(tb + p.SPT - 1, p.U, p.Gstream) - cta_of_unit(tb, p.U, p.Gstream) + 1);
    return ALS_NG * p.SPT;
}

template <int KB, bool WSTEP>
__global__ void __launch_bounds__(SOLVE_THREADS, (KB <= 2) ? (768 / SOLVE_THREADS) : ((KB >= 5 && KB <= 7) ? SOLVE_MINB_LARGE : 1)) als_solve_kernel(const HalfParams p) {
    using Cfg = SolveCfg<KB>;
    constexpr int KP = Cfg::KP, SU = Cfg::SU, XS = Cfg::XS, UPW = Cfg::UPW, LP = Cfg::LP, R = Cfg::R;
    pdl_prologue(WSTEP ? 5 : 3);
    const int stopped = *(volatile int *)&p.ctrl->stop;     // consumed after the loads below are in flight
    extern __shared__ __align__(16) unsigned char solve_smem[];
    double *sG = (double *)solve_smem;
    double *sM = sG + KP * KP;
    double *xs = sM + KP * KP;
    double *scr = xs + KP * XS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = p.k;

    for (int e = tid; e < KP * KP; e += SOLVE_THREADS) {
        sG[e] = p.G[e];
        sM[e] = p.Minv[e];
    }
    const long long unit0 = (long long)blockIdx.x * SU;
    const int tile = (int)(unit0 / TILE_UNITS), uoff = (int)(unit0 - (long long)tile * TILE_UNITS);
    const int nvalid_local = (int)((p.nvalid - unit0 < SU) ? (p.nvalid - unit0) : SU);
    // number of partial slots the stream kernel filled for this tile (see its flush())
    int expected = expected_slots(p, tile);
    if (p.shard && !WSTEP) expected = 1;                     // P = the all-reduced W'A, one slot per tile
    // (a) right-hand sides: sum the partials in slot order
    const bool peer_sum = !WSTEP && p.shard && p.peers != nullptr;
    for (int e = tid; e < SU * KP; e += SOLVE_THREADS) {
        const int u = e % SU, c = e / SU;
        double r = 0.0;
        if (peer_sum) {
            // W'A = sum over the ranks of their folded W_q'A_q, read over NVLink in rank order (all loads in flight)
            if (c < k && u < nvalid_local) {
                const ShardPeers *sp = p.peers;
                const size_t off = ((size_t)tile * KP + c) * (size_t)TILE_UNITS + uoff + u;
                const int world = sp->world, par = p.epoch & 1;
                double v[SHARD_MAXP];
                if (sp->push) {      // every rank has stored its buffer into slot q of THIS rank's block
                    const double *mine = shard_xchg(sp, sp->rank, par) + off;
                    const size_t stride = (size_t)sp->cnt;
#pragma unroll
                    for (int q = 0; q < SHARD_MAXP; ++q) v[q] = (q < world) ? ldcg(mine + (size_t)q * stride) : 0.0;
                } else {
#pragma unroll
                    for (int q = 0; q < SHARD_MAXP; ++q) v[q] = (q < world) ? ldcg(shard_xchg(sp, q, par) + off) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < SHARD_MAXP; ++q) r += v[q];
            }
        } else if (c < k && u < nvalid_local) {
            const double *Pt = p.P + ((size_t)tile * p.maxslots * KP + c) * (size_t)TILE_UNITS + uoff + u;
            if (expected <= 12) {                            // all slots in flight at once, summed in slot order
                double v[12];
#pragma unroll
                for (int s = 0; s < 12; ++s) v[s] = (s < expected) ? ldcg(Pt + (size_t)s * (KP * TILE_UNITS)) : 0.0;
#pragma unroll
                for (int s = 0; s < 12; ++s) r += v[s];
            } else {
#pragma unroll 4
                for (int s = 0; s < expected; ++s) r += ldcg(Pt + (size_t)s * (KP * TILE_UNITS));
            }
        }
        xs[c * XS + u] = r;
    }
    if (stopped) return;
    __syncthreads();
    // ---- from here on the warps of the CTA are independent ----
    // (b) exact NNLS, (c) write-back and statistics (lane = (unit, variables) as in nnls_warp)
    double d2 = 0.0, t2 = 0.0;
    unsigned long long nz = 0ull;
    double t1;
    if constexpr (KP <= 32)      // tableau in registers, one lane per row
        t1 = nnls_lanes<KP, XS>(sG, sM, k, xs, warp * UPW, UPW, nvalid_local, lane, scr + warp * NnlsCfg<KP>::SCR,
                                p.ctrl->count_draws ? &p.ctrl->draws : nullptr);
    else                         // compressed active-set solves, two variables per lane
        t1 = nnls_warp<KP, XS>(sG, sM, k, xs, scr + warp * NnlsCfg<KP>::SCR, warp * UPW, UPW, nvalid_local, lane);
    __syncwarp();
    {
        const int i = lane % LP;
        for (int uu = lane / LP; uu < UPW; uu += 32 / LP) {
            const int u = warp * UPW + uu;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int v = i + LP * q;
                if (v < k) {
                    const double x = xs[v * XS + u];
                    const size_t off = (size_t)v * p.ldout + (size_t)(unit0 + u);
                    if (u < nvalid_local) {
                        const double o = p.Fold[off];
                        d2 += (x - o) * (x - o);
                        if (x != 0.0) nz |= (1ull << v);
                        if (WSTEP) {
                            double gx = 0.0;
#pragma unroll 4
                            for (int j = 0; j < k; ++j) gx += sG[j * KP + v] * xs[j * XS + u];   // G symmetric: conflict-free
                            t2 += x * gx;
                        }
                    }
                    p.Fout[off] = x;
                }
            }
        }
    }
    d2 = warp_sum(d2);
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nz |= __shfl_xor_sync(0xffffffffu, nz, o);
    const int gw = (int)blockIdx.x * SOLVE_WARPS + warp;            // global warp index
    const int nwarps = (int)gridDim.x * SOLVE_WARPS;
    const int grp = gw / SOLVE_GROUP, gfirst = grp * SOLVE_GROUP;
    const int gsize = (nwarps - gfirst < SOLVE_GROUP) ? nwarps - gfirst : SOLVE_GROUP;
    const int ngroups = (nwarps + SOLVE_GROUP - 1) / SOLVE_GROUP;
    int last = 0;
    if (lane == 0) {
        p.spart[(size_t)gw * 4 + 0] = d2;
        p.spart[(size_t)gw * 4 + 1] = t1;
        p.spart[(size_t)gw * 4 + 2] = t2;
        if (!WSTEP) atomicOr(&p.ctrl->nzmask, nz);
        __threadfence();
        const int prev = atomicAdd(&p.group_counter[grp], 1);
        last = (prev == gsize - 1);
        if (last) p.group_counter[grp] = 0;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    // ---- level 1: this warp arrived last in its group: fold the group's warp partials in a fixed order ----
    {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int q = 0; q < SOLVE_GROUP / 32; ++q) {
            const int w = q * 32 + lane;
            if (w < gsize) {
                a0 += ldcg(&p.spart[(size_t)(gfirst + w) * 4 + 0]);
                a1 += ldcg(&p.spart[(size_t)(gfirst + w) * 4 + 1]);
                a2 += ldcg(&p.spart[(size_t)(gfirst + w) * 4 + 2]);
            }
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
        if (lane == 0) {
            p.spart2[(size_t)grp * 4 + 0] = a0;
            p.spart2[(size_t)grp * 4 + 1] = a1;
            p.spart2[(size_t)grp * 4 + 2] = a2;
            __threadfence();
            const int prev = atomicAdd(&p.ctrl->groups_done[WSTEP ? 1 : 0], 1);
            last = (prev == ngroups - 1);
        }
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    // ---- level 2 (one warp, once per half-step): iteration scalars and the control block ----
    double sd2 = 0.0, st1 = 0.0, st2 = 0.0;
    for (int g = lane; g < ngroups; g += 32) {
        sd2 += ldcg(&p.spart2[(size_t)g * 4 + 0]);
        st1 += ldcg(&p.spart2[(size_t)g * 4 + 1]);
        st2 += ldcg(&p.spart2[(size_t)g * 4 + 2]);
    }
    sd2 = warp_sum(sd2); st1 = warp_sum(st1); st2 = warp_sum(st2);
    if (lane == 0) {
        AlsCtrl *c = p.ctrl;
        c->groups_done[WSTEP ? 1 : 0] = 0;
        if (!WSTEP) {
            c->dH2 = sd2;
            const unsigned long long fullm = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
            const unsigned long long m = c->nzmask;
            c->nzmask = 0ull;
            if ((m & fullm) != fullm) { c->stop = 1; c->reason = 2; }
        } else if (p.shard) {
            // this rank's share of the sums; the host all-reduces them and als_shard_finish_kernel closes the iteration
            if (p.peers) {
                const ShardPeers *sp = p.peers;
                double *xs2 = shard_xscal(sp, sp->rank, p.epoch & 1);
                xs2[0] = sd2; xs2[1] = st1; xs2[2] = st2; xs2[3] = 0.0;
            } else {
                p.xscal[0] = sd2; p.xscal[1] = st1; p.xscal[2] = st2; p.xscal[3] = 0.0;
            }
        } else {
            finish_iteration(c, p.trace, sd2, st1, st2);
        }
        __threadfence();
    }
    if (WSTEP && p.shard && p.peers) {
        // peer mode: publish this rank's scalars and leave.  The iteration is closed by als_shard_close_kernel, which the
        // host places after the NEXT H-step's stream kernel (or at the end of the batch): by then every rank's flag has
        // long arrived, so the skew between the ranks' last warps is absorbed by 140 us of streaming instead of sitting
        // on the critical path (+23 us per iteration at eight GPUs).
        // (the flag itself -- a system-scope fence and `world` remote stores -- is raised by the next kernel that has
        // time for it: the side job of the next H-step's stream kernel, or the kernel that closes the batch)
        if (lane == 0) {
            p.ctrl->pending_epoch = p.epoch;
            p.ctrl->unpublished_epoch = p.epoch;
            __threadfence();
        }
    }
}

// Raises the flag of the W-step scalars this rank wrote last (flagsB) in every rank's block, if nobody has yet.  Called by
// thread t of a group of at least `world` threads of ONE block per launch; the W-step solve kernel that wrote the
// scalars is an earlier launch of the same stream.
__device__ __forceinline__ void shard_publish_scalars(const HalfParams &p, int t) {
    const int e = *(volatile int *)&p.ctrl->unpublished_epoch;
    if (e == 0 || p.peers == nullptr) return;
    shard_publish(p.peers, SHARD_OFF_FLAGSB, e, t);
}

// the end of R's loop body for the iteration whose W-step scalars were written last (one warp, peer-memory sessions):
// waits for every rank's flag, adds the ranks' sums in rank order (bit-identical on every rank)
__device__ __forceinline__ void shard_close_iteration(const HalfParams &p, int lane) {
    AlsCtrl *c = p.ctrl;
    const int e = *(volatile int *)&c->pending_epoch;
    if (e == 0 || p.peers == nullptr) return;
    const ShardPeers *sp = p.peers;
    shard_publish_scalars(p, lane);                       // (end of a batch: no stream kernel has done it)
    __syncwarp();
    if (lane < sp->world) shard_wait_flag((const int *)(sp->base[sp->rank] + SHARD_OFF_FLAGSB) + lane, e, lane, c);
    __syncwarp();
    // lane q fetches rank q's three sums (all remote loads in flight at once); lane 0 adds them in rank order
    double x0 = 0.0, x1 = 0.0, x2 = 0.0;
    if (lane < sp->world) {
        const double *xq = shard_xscal(sp, lane, e & 1);
        x0 = ldcg(xq + 0); x1 = ldcg(xq + 1); x2 = ldcg(xq + 2);
    }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int q = 0; q < sp->world; ++q) {
        a0 += __shfl_sync(0xffffffffu, x0, q);
        a1 += __shfl_sync(0xffffffffu, x1, q);
        a2 += __shfl_sync(0xffffffffu, x2, q);
    }
    if (lane == 0) {
        if (c->reason != 5) finish_iteration(c, p.trace, a0, a1, a2);      // (5: a peer never arrived)
        c->pending_epoch = 0;
        c->unpublished_epoch = 0;
        __threadfence();
    }
}

// Closes the iteration whose W-step scalars were published last (peer-memory sessions): waits for every rank's flag,
// adds the ranks' sums in rank order (bit-identical on every rank) and runs the end of R's loop body.  One warp.  It
// runs between the stream kernel of the next H-step and its fold kernel: every kernel after it sees the decision, the
// stream kernel before it has at worst streamed A once for an iteration R would not have started (it writes partials
// and the Gram of W only; if it found that Gram singular, the convergence found here takes precedence, as in R).
__global__ void __launch_bounds__(32) als_shard_close_kernel(const HalfParams p) {
    pdl_prologue(1);
    shard_close_iteration(p, threadIdx.x);
}

// ---------------------------------------------------------------------------
// side job of the stream kernel: Gram matrix of the FIXED factor and its inverse
// ---------------------------------------------------------------------------
// Run by the three idle warps of the producer warpgroup (96 threads, named barrier 2) of every stream CTA while the
// other warps stream A: CTA b forms the partial Gram of its slice of the fixed factor's units, the last CTA to
// arrive folds the 148 partials in CTA order and inverts the k x k matrix in shared memory.  The solve kernel of the
// same half-step reads G and G^-1; nothing on the critical path waits for them.  R's solve() failure ("system is
// computationally singular") is raised here: reason 3 = W'W (H-step), 4 = HH' (W-step).
#define GRAM_THREADS 96
// development aid (MFB_GRAM_TIMING): globaltimer stamps of the side job of the CTA that ends up inverting, and of its consumers
__device__ unsigned long long g_gram_stamps[16];
#define GRAM_STAMP(i) do { if (t == 0) g_gram_stamps[(i)] = globaltimer_ns(); } while (0)
#define GRAM_GROUP 16            // stream CTAs per fold group
__device__ __forceinline__ void gram_bar() { asm volatile("bar.sync 2, 96;" ::: "memory"); }

// Partial Gram of the units [lo, hi) of the fixed factor on the FP64 tensor cores.  The KB (KB + 1) / 2 blocks of 8 x 8 on
// and above the diagonal are dealt round-robin to the three warps (W = this warp's share, all indices compile-time);
// every warp walks ALL the units once, 4 per DMMA, with the fragments of two k-steps in flight -- fragment b of a
// k-step, F[8b + g][u + tt], is the A operand of the blocks in block row b and the B operand of those in block column b.
// The block below the diagonal is the transpose: the same products in the same order, copied.
template <int KB, int W>
__device__ __forceinline__ void gram_partial_blocks(const HalfParams &p, long long lo, long long hi, int lane, double *sGm) {
    constexpr int KP = KB * 8, NB = KB * (KB + 1) / 2;
    constexpr int MY = (NB > W) ? (NB - W + 2) / 3 : 0;
    if (MY == 0) return;
    const int g = lane >> 2, tt = lane & 3;
    double acc[MY > 0 ? MY : 1][2];
#pragma unroll
    for (int q = 0; q < MY; ++q) acc[q][0] = acc[q][1] = 0.0;
    const double *fp = p.Fin + (size_t)g * p.ldin + tt;
    for (long long u = lo; u < hi; u += 8) {
        double fr[2][KB];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            const long long uu = u + 4 * s2;
            const bool ok = uu + tt < hi;
#pragma unroll
            for (int bq = 0; bq < KB; ++bq) fr[s2][bq] = ok ? fp[(size_t)(8 * bq) * p.ldin + uu] : 0.0;
        }
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            int idx = 0, q = 0;
#pragma unroll
            for (int i = 0; i < KB; ++i)
#pragma unroll
                for (int j = i; j < KB; ++j) {
                    if (idx % 3 == W) {
                        dmma884(acc[q][0], acc[q][1], fr[s2][i], fr[s2][j]);
                        ++q;
                    }
                    ++idx;
                }
        }
    }
    {
        int idx = 0, q = 0;
#pragma unroll
        for (int i = 0; i < KB; ++i)
#pragma unroll
            for (int j = i; j < KB; ++j) {
                if (idx % 3 == W) {
                    const int row = 8 * i + g, col = 8 * j + 2 * tt;
                    sGm[row * KP + col] = acc[q][0];
                    sGm[row * KP + col + 1] = acc[q][1];
                    sGm[col * KP + row] = acc[q][0];
                    sGm[(col + 1) * KP + row] = acc[q][1];
                    ++q;
                }
                ++idx;
            }
    }
}

// Fixed-order sum of `nsrc` symmetric KP x KP matrices (src + q * KP * KP, q = 0 .. nsrc - 1) by the 96 side-job threads:
// only the 8 x 8 blocks on and above the diagonal are read (pairs of entries, 16-byte loads, eight sources in flight);
// f(row, col, v0, v1) receives the sums of the entries (row, col) and (row, col + 1) and mirrors them itself.
template <int KB, typename F>
__device__ __forceinline__ void gram_fold_upper(const double *src, int nsrc, int t, F f) {
    constexpr int KP = KB * 8, NB = KB * (KB + 1) / 2;
    for (int q = t; q < NB * 32; q += GRAM_THREADS) {
        int bidx = q >> 5, i = 0;
        while (bidx >= KB - i) { bidx -= KB - i; ++i; }
        const int j = i + bidx, rem = q & 31;
        const int row = 8 * i + (rem >> 2), col = 8 * j + 2 * (rem & 3);
        const double *sp = src + row * KP + col;
        double s0 = 0.0, s1 = 0.0;
        for (int q0 = 0; q0 < nsrc; q0 += 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = (q0 + u < nsrc) ? __ldcg((const double2 *)(sp + (size_t)(q0 + u) * (KP * KP))) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 8; ++u) { s0 += v[u].x; s1 += v[u].y; }
        }
        f(row, col, s0, s1);
    }
}

template <int KB, bool WSTEP>
__device__ __noinline__ void gram_side_job(const HalfParams &p, const int t, double *sGm, int *gflag) {
    constexpr int KP = KB * 8;
    const int k = p.k, G = (int)gridDim.x, b = (int)blockIdx.x;
    const int w = t >> 5, lane = t & 31;
    // peer-memory sessions: the flag of the previous W-step's scalars (written by its solve kernel's last warp) is
    // raised here, where nobody waits for the fence
    if (!WSTEP && p.shard && p.peers && b == 0) {
        const int e = *(volatile int *)&p.ctrl->unpublished_epoch;      // (uniform: cleared only after the barrier)
        if (e != 0) {
            shard_publish(p.peers, SHARD_OFF_FLAGSB, e, t);
            gram_bar();
            if (t == 0) p.ctrl->unpublished_epoch = 0;
        }
    }
    // --- partial Gram of this CTA's slice of units ---
    const unsigned long long t_begin = globaltimer_ns();
    const long long lo = (long long)b * p.nfix / G, hi = (long long)(b + 1) * p.nfix / G;
    if (w == 0) gram_partial_blocks<KB, 0>(p, lo, hi, lane, sGm);
    else if (w == 1) gram_partial_blocks<KB, 1>(p, lo, hi, lane, sGm);
    else gram_partial_blocks<KB, 2>(p, lo, hi, lane, sGm);
    gram_bar();
    const unsigned long long t_part = globaltimer_ns();
    for (int e = t; e < KP * KP; e += GRAM_THREADS) p.gram_part[(size_t)b * (KP * KP) + e] = sGm[e];
    __threadfence();
    gram_bar();
    // --- level 1: the last CTA of a group of GRAM_GROUP folds the group (CTA order) ---
    const int grp = b / GRAM_GROUP, gfirst = grp * GRAM_GROUP;
    const int gsize = (G - gfirst < GRAM_GROUP) ? G - gfirst : GRAM_GROUP;
    const int ngroups = (G + GRAM_GROUP - 1) / GRAM_GROUP;
    if (t == 0) {
        const int prev = atomicAdd(&p.gram_counter[1 + grp], 1);
        const int lastone = (prev == gsize - 1);
        if (lastone) p.gram_counter[1 + grp] = 0;
        *gflag = lastone;
    }
    gram_bar();
    if (!*gflag) return;
    const unsigned long long t_l1 = globaltimer_ns();
    __threadfence();
    {
        double *dst = p.gram_part2 + (size_t)grp * (KP * KP);
        gram_fold_upper<KB>(p.gram_part + (size_t)gfirst * (KP * KP), gsize, t, [&](int row, int col, double v0, double v1) {
            dst[row * KP + col] = v0;
            dst[row * KP + col + 1] = v1;
            dst[col * KP + row] = v0;
            dst[(col + 1) * KP + row] = v1;
        });
    }
    __threadfence();
    gram_bar();
    if (t == 0) {
        const int prev = atomicAdd(&p.gram_counter[0], 1);
        const int lastone = (prev == ngroups - 1);
        if (lastone) p.gram_counter[0] = 0;
        *gflag = lastone;
    }
    gram_bar();
    if (!*gflag) return;
    __threadfence();
    // --- level 2 (one CTA per launch): fold the groups, publish G, invert ---
    if (t == 0) { g_gram_stamps[0] = t_begin; g_gram_stamps[1] = t_part; g_gram_stamps[2] = t_l1; }
    GRAM_STAMP(3);
    const double reg_all = p.reg_all, reg_diag = p.reg_diag;    // 0 for als_nmf (x + 0.0 is exact)
    gram_fold_upper<KB>(p.gram_part2, ngroups, t, [&](int row, int col, double v0, double v1) {
        v0 += reg_all + (row == col ? reg_diag : 0.0);
        v1 += reg_all + (row == col + 1 ? reg_diag : 0.0);
        if (!(row < k && col < k)) v0 = 0.0;
        if (!(row < k && col + 1 < k)) v1 = 0.0;
        sGm[row * KP + col] = v0;      p.G[row * KP + col] = v0;
        sGm[row * KP + col + 1] = v1;  p.G[row * KP + col + 1] = v1;
        sGm[col * KP + row] = v0;      p.G[col * KP + row] = v0;
        sGm[(col + 1) * KP + row] = v1; p.G[(col + 1) * KP + row] = v1;
    });
    gram_bar();
    // large ranks: the inversion is a long serial chain that competes with this CTA's own DMMA warps for issue slots
    // (k = 64: ~350 us on the side-job warps, the tail of the whole launch); a one-block kernel does it in ~15 us
    if (p.invert_later) return;
    // row-sharded H-step: this is only this rank's W_r'W_r
    if (p.shard && !WSTEP) {
        if (!p.peers) return;            // callback exchange: summed with W_r'A_r by the host, inverted by als_shard_invert_kernel
        // peer memory: publish, wait for the other ranks' Gram matrices, add them in rank order -- all of it (and the
        // inversion below) while the consumer warps are still streaming A
        const ShardPeers *sp = p.peers;
        const int par = p.epoch & 1;
        double *mine = shard_gram(sp, sp->rank, par);
        for (int e = t; e < KP * KP; e += GRAM_THREADS) mine[e] = sGm[e];
        gram_bar();
        shard_publish(sp, SHARD_OFF_FLAGSG, p.epoch, t);
        if (p.peer_invert_later) return;     // summed and inverted after the stream kernel (als_shard_close_invert_kernel)
        if (t < sp->world) shard_wait_flag((const int *)(sp->base[sp->rank] + SHARD_OFF_FLAGSG) + t, p.epoch, t, p.ctrl);
        gram_bar();
        if (*(volatile int *)&p.ctrl->stop) return;          // a peer never arrived
        for (int e = t; e < KP * KP; e += GRAM_THREADS) {
            const int c1 = e / KP, c2 = e - c1 * KP;
            double v = 0.0;
            if (c1 < k && c2 < k) {
                double g[SHARD_MAXP];                        // all ranks' loads in flight, then added in rank order
#pragma unroll
                for (int q = 0; q < SHARD_MAXP; ++q) g[q] = (q < sp->world) ? ldcg(shard_gram(sp, q, par) + e) : 0.0;
#pragma unroll
                for (int q = 0; q < SHARD_MAXP; ++q) v += g[q];
            }
            sGm[e] = v;
            p.G[e] = v;
        }
        gram_bar();
    }
    // all 96 threads on the shared matrix (two barriers per pivot): a single warp with the rows in registers is
    // latency-bound (~35 us at k = 16 and worse next to the streaming warps), which small or sharded matrices feel
    GRAM_STAMP(4);
    const int sing = gauss_jordan_inverse(sGm, k, KP, t, GRAM_THREADS, sGm + KP * KP, [] { gram_bar(); });
    GRAM_STAMP(5);
    for (int e = t; e < KP * KP; e += GRAM_THREADS) {
        const int c1 = e / KP, c2 = e - c1 * KP;
        p.Minv[e] = (c1 < k && c2 < k) ? sGm[e] : 0.0;
    }
    if (t == 0 && sing) {
        p.ctrl->reason = WSTEP ? 4 : 3;
        __threadfence();
        p.ctrl->stop = 1;
    }
    __threadfence();
}

// F32 = the resident matrix is stored in single precision (half the HBM traffic): a stage then covers 32 rows
// (H-step) / a 128 x 32 tile of floats (W-step), the fragments are widened to FP64 in registers and everything
// downstream (FP64 tensor-core accumulation, partials, factors, NNLS) is unchanged.
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int KB, bool WSTEP, bool F32>
__global__ void __launch_bounds__(ALS_THREADS, 1)
als_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmF, const HalfParams p,
                  const int nstage) {
    constexpr int KP = KB * 8;
    constexpr int A_BYTES = (WSTEP && !F32) ? 32768 : 16384;   // A box(es) of one stage
    constexpr int F_BOX = KP * 128;                      // one 16-unit box of the fixed factor: [KP][16] doubles
    constexpr int NFB = (WSTEP || F32) ? 2 : 1;          // factor boxes per stage (32 columns / 32 or 16 rows of reduction)
    constexpr int STAGE_BYTES = A_BYTES + NFB * F_BOX;
    pdl_prologue(WSTEP ? 4 : 0);
    if (*(volatile int *)&p.ctrl->stop) return;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: ring | barriers
    unsigned char *ring = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = (uint64_t *)(ring + (size_t)nstage * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + nstage;
    double *sGm = (double *)(empty_bar + nstage);        // KP*KP + 2 doubles: Gram side job
    int *gflag = (int *)(sGm + KP * KP + 2);
    const uint32_t ring_s = smem_u32(ring);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const long long u0 = (long long)blockIdx.x * p.U / G;
    const long long u1 = (long long)(blockIdx.x + 1) * p.U / G;

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const bool has_work = u0 < u1;
    const int tile0 = (int)(u0 / p.SPT), st0 = (int)(u0 - (long long)tile0 * p.SPT);
    const int nunits = (int)(u1 - u0);

    if (warp >= 4 * ALS_NG) {
        // ===== warpgroup 2: TMA producer (warp 8) + Gram side job (warps 9-11); hands registers to the consumers =====
#if ALS_NG == 2
        if (has_work) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
#endif
        if (warp > 4 * ALS_NG) {
            gram_side_job<KB, WSTEP>(p, tid - (ALS_CONSUMERS + 32), sGm, gflag);
        } else if (has_work && lane == 0) {
            int stage = 0, tile = tile0, st = st0;
            uint32_t phase = 0;
            for (int it = 0; it < nunits; ++it) {
                mbar_wait(&empty_bar[stage], phase ^ 1, 1, it, stage);
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                unsigned char *dst = ring + (size_t)stage * STAGE_BYTES;
                if (!WSTEP && !F32) {
                    tma_load_2d(dst, &tmA, st * 16, tile * TILE_UNITS, &full_bar[stage]);
                    tma_load_2d(dst + A_BYTES, &tmF, st * 16, 0, &full_bar[stage]);
                } else if (!WSTEP) {
                    // 32 rows of floats per column = the same 128-byte smem rows; two 16-row slabs of W
                    tma_load_2d(dst, &tmA, st * 32, tile * TILE_UNITS, &full_bar[stage]);
                    tma_load_2d(dst + A_BYTES, &tmF, st * 32, 0, &full_bar[stage]);
                    tma_load_2d(dst + A_BYTES + F_BOX, &tmF, st * 32 + 16, 0, &full_bar[stage]);
                } else {
                    // one 3-D box {16 rows, 32 columns, 8 row groups} (FP32: {32 rows, 32 columns, 4 row groups}):
                    // lands as [row group][column][128 bytes of rows]
                    tma_load_3d(dst, &tmA, 0, st * 32, tile * (F32 ? 4 : 8), &full_bar[stage]);
                    tma_load_2d(dst + A_BYTES, &tmF, st * 32, 0, &full_bar[stage]);
                    tma_load_2d(dst + A_BYTES + F_BOX, &tmF, st * 32 + 16, 0, &full_bar[stage]);
                }
                if (++stage == nstage) { stage = 0; phase ^= 1; }
                if (++st == p.SPT) { st = 0; ++tile; }
            }
        }
        return;
    }

    // ===================== consumers: NG groups of 4 warps (one warp per SM sub-partition) =====================
    // Group `grp` owns the units it == grp (mod NG) of this CTA's range, so the two warps that share a
    // sub-partition are one stage out of phase: one issues DMMAs while the other waits / loads fragments.
    if (!has_work) return;
#if ALS_NG == 2
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
#endif
    constexpr int NG = ALS_NG;
    const int grp = warp >> 2, wg = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    double acc[KB][4][2];                        // [rank block][v (H-step: column block) or 2*rg+s (W-step: row)][2]
#pragma unroll
    for (int cb = 0; cb < KB; ++cb)
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[cb][a][0] = acc[cb][a][1] = 0.0;

    auto flush = [&](int tile) {
        const long long tb = (long long)tile * p.SPT;
        int slot;
        if (p.U >= G)            // every CTA owns >= 1 unit: contributors to a tile are contiguous CTAs
            slot = NG * ((int)blockIdx.x - cta_of_unit(tb, p.U, G)) + grp;   // one partial per consumer group of each CTA
        else                     // fewer units than CTAs: each non-empty CTA owns exactly one unit
            slot = NG * (int)(u0 - tb) + grp;
        double *Pb = p.P + ((size_t)tile * p.maxslots + slot) * (size_t)(KP * TILE_UNITS);
#pragma unroll
        for (int cb = 0; cb < KB; ++cb) {
            if (!WSTEP) {
                // acc[cb][v] = R[8cb+g][32wg + 8v + 2t + {0,1}]
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    *(double2 *)&Pb[(size_t)(8 * cb + g) * TILE_UNITS + 32 * wg + 8 * v + 2 * t] =
                        make_double2(acc[cb][v][0], acc[cb][v][1]);
            } else if (!F32) {
                // acc[cb][2rg+s][e] = R[row 32wg + 16rg + 2g + s][c = 8cb + t + 4e]   (rank index permuted, see the W-step loop)
#pragma unroll
                for (int rg = 0; rg < 2; ++rg)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        *(double2 *)&Pb[(size_t)(8 * cb + t + 4 * e) * TILE_UNITS + 32 * wg + 16 * rg + 2 * g] =
                            make_double2(acc[cb][2 * rg][e], acc[cb][2 * rg + 1][e]);
            } else {
                // FP32 tiles: acc[cb][s][e] = R[row 32wg + 4g + s][c = 8cb + t + 4e]
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        *(double2 *)&Pb[(size_t)(8 * cb + t + 4 * e) * TILE_UNITS + 32 * wg + 4 * g + 2 * s2] =
                            make_double2(acc[cb][2 * s2][e], acc[cb][2 * s2 + 1][e]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[cb][a][0] = acc[cb][a][1] = 0.0;
        }
    };

    // Main loop.  One stage = one unit: an A box plus the matching 16- (H-step) or 32-wide (W-step) slab of
    // the fixed factor, both landed by TMA with SWIZZLE_128B: a "row" of a box is 128 bytes (16 doubles along
    // the contiguous direction) whose 16-byte chunk c sits at chunk c ^ (row & 7).  Every fragment read below
    // is a 16-byte ld.shared whose quarter-warp (g in {2a, 2a+1}, t = 0..3) touches 8 distinct chunks.
    const int nst_r = pin_reg(nstage), spt_r = pin_reg(p.SPT);
    const int pg = (g >> 1) | ((g & 1) << 2);
    // per-lane byte offsets inside a stage (everything else is an immediate)
    uint32_t offA, offF;
    if (!WSTEP) {
        offA = (uint32_t)((32 * wg + g) * 128 + (((2 * t) ^ g) << 4));      // column 32wg + g (+8v), chunk 2t (^1 for rows 4t+2..)
        offF = F32 ? (uint32_t)(A_BYTES + (t >> 1) * F_BOX + g * 128)       // FP32: rows 8t..8t+7 = slab t>>1, chunks 4(t&1)+e
                   : (uint32_t)(A_BYTES + g * 128 + (((2 * t) ^ g) << 4));  // rank g (+8cb)
    } else if (F32) {
        offA = (uint32_t)(wg * 4096 + (2 * t) * 128 + ((g ^ (2 * t)) << 4));       // row group wg, column 2t (+e +8q), chunk g = rows 4g..4g+3
        offF = (uint32_t)(A_BYTES + pg * 128 + ((t ^ pg) << 4));
    } else {
        offA = (uint32_t)(2 * wg * 4096 + (2 * t) * 128 + ((g ^ (2 * t)) << 4));   // row group 2wg (+rg), column 2t (+e +8q), chunk g
        offF = (uint32_t)(A_BYTES + pg * 128 + ((t ^ pg) << 4));                   // rank pi(g) (+8cb), chunk t (+4(q&1)) ^ pi(g)
    }
    int tile = tile0, st = st0;
    int stage = grp;                     // this group's first unit sits in stage `grp`
    uint32_t phase = 0;
    for (int it = 0; it < nunits; ++it) {
        if ((it % NG) == grp) {
            const uint32_t sb = ring_s + (uint32_t)stage * STAGE_BYTES;
            mbar_wait(&full_bar[stage], phase, 2, it, stage);
            if (!WSTEP && F32) {
                // stage = 32 rows x 128 columns of floats; lane (g, t) owns rows 8t..8t+7 of the columns g (+8v) of its
                // 32-column slice: the same two 16-byte chunks per column as in FP64, four rows each
                float4 b4[4][2];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    b4[v][0] = lds128f(sb + offA + v * 1024);
                    b4[v][1] = lds128f(sb + (offA ^ 16) + v * 1024);
                }
#pragma unroll
                for (int cb = 0; cb < KB; ++cb) {
                    double2 f2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        f2[e] = lds128(sb + offF + cb * 1024 + ((((t & 1) * 4 + e) ^ g) << 4));   // rows 8t+2e, 8t+2e+1 of rank 8cb+g
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 q4 = b4[v][e >> 1];
                            const double x0 = (double)((e & 1) ? q4.z : q4.x), x1 = (double)((e & 1) ? q4.w : q4.y);
                            dmma884(acc[cb][v][0], acc[cb][v][1], f2[e].x, x0);
                            dmma884(acc[cb][v][0], acc[cb][v][1], f2[e].y, x1);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            } else if (WSTEP && F32) {
                // stage = 128 rows x 32 columns of floats as [row group of 32][column][32 rows]; lane (g, t) owns rows
                // 4g..4g+3 of its warp's row group at the columns 8q + 2t + e
                double2 f[KB][4];
#pragma unroll
                for (int cb = 0; cb < KB; ++cb)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        f[cb][q] = lds128(sb + (q >> 1) * F_BOX + ((offF + cb * 1024) ^ ((q & 1) << 6)));
                float4 a4[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    a4[q][0] = lds128f(sb + offA + q * 1024);
                    a4[q][1] = lds128f(sb + ((offA + 128) ^ 16) + q * 1024);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int cb = 0; cb < KB; ++cb) {
                        dmma884(acc[cb][0][0], acc[cb][0][1], (double)a4[q][0].x, f[cb][q].x);
                        dmma884(acc[cb][1][0], acc[cb][1][1], (double)a4[q][0].y, f[cb][q].x);
                        dmma884(acc[cb][2][0], acc[cb][2][1], (double)a4[q][0].z, f[cb][q].x);
                        dmma884(acc[cb][3][0], acc[cb][3][1], (double)a4[q][0].w, f[cb][q].x);
                    }
#pragma unroll
                    for (int cb = 0; cb < KB; ++cb) {
                        dmma884(acc[cb][0][0], acc[cb][0][1], (double)a4[q][1].x, f[cb][q].y);
                        dmma884(acc[cb][1][0], acc[cb][1][1], (double)a4[q][1].y, f[cb][q].y);
                        dmma884(acc[cb][2][0], acc[cb][2][1], (double)a4[q][1].z, f[cb][q].y);
                        dmma884(acc[cb][3][0], acc[cb][3][1], (double)a4[q][1].w, f[cb][q].y);
                    }
                }
            } else if (!WSTEP && KB >= 6) {
                // large ranks: the accumulator tile alone is 2 * KB * 8 registers, so the A fragments are taken two column
                // blocks at a time and the factor fragments rank block by rank block (same accumulation order per
                // accumulator as below; nothing spills)
#pragma unroll
                for (int vh = 0; vh < 2; ++vh) {
                    double2 b[2][2];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        b[v][0] = lds128(sb + offA + (2 * vh + v) * 1024);
                        b[v][1] = lds128(sb + (offA ^ 16) + (2 * vh + v) * 1024);
                    }
#pragma unroll
                    for (int cb = 0; cb < KB; ++cb) {
                        const double2 f0 = lds128(sb + offF + cb * 1024), f1 = lds128(sb + (offF ^ 16) + cb * 1024);
#pragma unroll
                        for (int v = 0; v < 2; ++v) dmma884(acc[cb][2 * vh + v][0], acc[cb][2 * vh + v][1], f0.x, b[v][0].x);
#pragma unroll
                        for (int v = 0; v < 2; ++v) dmma884(acc[cb][2 * vh + v][0], acc[cb][2 * vh + v][1], f0.y, b[v][0].y);
#pragma unroll
                        for (int v = 0; v < 2; ++v) dmma884(acc[cb][2 * vh + v][0], acc[cb][2 * vh + v][1], f1.x, b[v][1].x);
#pragma unroll
                        for (int v = 0; v < 2; ++v) dmma884(acc[cb][2 * vh + v][0], acc[cb][2 * vh + v][1], f1.y, b[v][1].y);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            } else if (WSTEP && KB >= 6) {
#pragma unroll
                for (int rg = 0; rg < 2; ++rg) {
                    double2 a[4][2];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        a[q][0] = lds128(sb + offA + rg * 4096 + q * 1024);
                        a[q][1] = lds128(sb + ((offA + 128) ^ 16) + rg * 4096 + q * 1024);
                    }
#pragma unroll
                    for (int cb = 0; cb < KB; ++cb) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double2 fq = lds128(sb + (q >> 1) * F_BOX + ((offF + cb * 1024) ^ ((q & 1) << 6)));
                            dmma884(acc[cb][2 * rg][0], acc[cb][2 * rg][1], a[q][0].x, fq.x);
                            dmma884(acc[cb][2 * rg + 1][0], acc[cb][2 * rg + 1][1], a[q][0].y, fq.x);
                            dmma884(acc[cb][2 * rg][0], acc[cb][2 * rg][1], a[q][1].x, fq.y);
                            dmma884(acc[cb][2 * rg + 1][0], acc[cb][2 * rg + 1][1], a[q][1].y, fq.y);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            } else if (!WSTEP) {
                // m8n8k4: A-operand = W fragment [M = rank 8cb+g][K = row 4t+s], B-operand = A fragment [K = row 4t+s][N = column g]
                double2 b[4][2], f[KB][2];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    b[v][0] = lds128(sb + offA + v * 1024);
                    b[v][1] = lds128(sb + (offA ^ 16) + v * 1024);
                }
#pragma unroll
                for (int cb = 0; cb < KB; ++cb) {
                    f[cb][0] = lds128(sb + offF + cb * 1024);
                    f[cb][1] = lds128(sb + (offF ^ 16) + cb * 1024);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int cb = 0; cb < KB; ++cb)
#pragma unroll
                        for (int v = 0; v < 4; ++v) dmma884(acc[cb][v][0], acc[cb][v][1], f[cb][h].x, b[v][h].x);
#pragma unroll
                    for (int cb = 0; cb < KB; ++cb)
#pragma unroll
                        for (int v = 0; v < 4; ++v) dmma884(acc[cb][v][0], acc[cb][v][1], f[cb][h].y, b[v][h].y);
                }
            } else {
                // m8n8k4: A-operand = A fragment [M = row 2g+s][K = column j(t)], B-operand = H fragment [K = column j(t)][N = rank pi(g)]
                // with j(t) = 8q + 2t + e and pi(g) = (g >> 1) | ((g & 1) << 2), so that D column 2t+e is rank t + 4e.
                double2 f[KB][4];
#pragma unroll
                for (int cb = 0; cb < KB; ++cb)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        f[cb][q] = lds128(sb + (q >> 1) * F_BOX + ((offF + cb * 1024) ^ ((q & 1) << 6)));
#pragma unroll
                for (int rg = 0; rg < 2; ++rg) {
                    double2 a[4][2];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        a[q][0] = lds128(sb + offA + rg * 4096 + q * 1024);
                        a[q][1] = lds128(sb + ((offA + 128) ^ 16) + rg * 4096 + q * 1024);
                    }
                    if (rg == 1) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#pragma unroll
                        for (int cb = 0; cb < KB; ++cb) {
                            dmma884(acc[cb][2 * rg][0], acc[cb][2 * rg][1], a[q][0].x, f[cb][q].x);
                            dmma884(acc[cb][2 * rg + 1][0], acc[cb][2 * rg + 1][1], a[q][0].y, f[cb][q].x);
                        }
#pragma unroll
                        for (int cb = 0; cb < KB; ++cb) {
                            dmma884(acc[cb][2 * rg][0], acc[cb][2 * rg][1], a[q][1].x, f[cb][q].y);
                            dmma884(acc[cb][2 * rg + 1][0], acc[cb][2 * rg + 1][1], a[q][1].y, f[cb][q].y);
                        }
                    }
                }
            }
            stage += NG;
            if (stage >= nst_r) { stage -= nst_r; phase ^= 1; }
        }
        const bool last_of_tile = (st + 1 == spt_r) || (it + 1 == nunits);
        if (last_of_tile) flush(tile);
        if (++st == spt_r) { st = 0; ++tile; }
    }
}

#include "als_tf32.cuh"

// ---------------------------------------------------------------------------
// stand-alone Gram + inverse of a factor [KP][ld] over `len` entries (start / restart only)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_partial_kernel(const double *__restrict__ F, int64_t ld, int64_t len,
                                                           int k, int KP, double *__restrict__ part) {
    extern __shared__ double sh[];   // [k][65]
    const int nb = gridDim.x;
    const int64_t per = (len + nb - 1) / nb;
    const int64_t lo = blockIdx.x * per, hi = (lo + per < len) ? lo + per : len;
    const int npairs = k * k;
    double accp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) accp[i] = 0.0;
    for (int64_t base = lo; base < hi; base += 64) {
        const int cnt = (int)((hi - base < 64) ? hi - base : 64);
        for (int idx = threadIdx.x; idx < k * 64; idx += 256) {
            const int c = idx >> 6, u = idx & 63;
            sh[c * 65 + u] = (u < cnt) ? F[c * ld + base + u] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int e = threadIdx.x + q * 256;
            if (e < npairs) {
                const int c1 = e / k, c2 = e - c1 * k;
                double s = 0.0;
                for (int u = 0; u < 64; ++u) s += sh[c1 * 65 + u] * sh[c2 * 65 + u];
                accp[q] += s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int e = threadIdx.x + q * 256;
        if (e < npairs) {
            const int c1 = e / k, c2 = e - c1 * k;
            part[(size_t)blockIdx.x * KP * KP + c1 * KP + c2] = accp[q];
        }
    }
}

__global__ void __launch_bounds__(256) gram_finalize_kernel(const double *__restrict__ part, int nb, int k, int KP,
                                                            double *__restrict__ Gout, double *__restrict__ Minvout,
                                                            int *__restrict__ singular) {
    extern __shared__ double sh[];   // KP*KP + 2
    double *sG = sh, *scr = sh + KP * KP;
    for (int e = threadIdx.x; e < KP * KP; e += 256) {
        const int c1 = e / KP, c2 = e - c1 * KP;
        double s = 0.0;
        if (c1 < k && c2 < k)
            for (int b = 0; b < nb; ++b) s += part[(size_t)b * KP * KP + e];
        sG[e] = s;
        Gout[e] = s;
    }
    __syncthreads();
    const int sing = gauss_jordan_inverse(sG, k, KP, threadIdx.x, 256, scr, [] { __syncthreads(); });
    for (int e = threadIdx.x; e < KP * KP; e += 256) {
        const int c1 = e / KP, c2 = e - c1 * KP;
        Minvout[e] = (c1 < k && c2 < k) ? sG[e] : 0.0;
    }
    if (threadIdx.x == 0) *singular = sing;
}

// ---------------------------------------------------------------------------
// row-sharded ALS: A and W are split by rows over the ranks, H is replicated
// ---------------------------------------------------------------------------
// The H-step needs  W'A = sum_r W_r'A_r  and  W'W = sum_r W_r'W_r : after the local stream kernel this kernel sums the
// stream-K partials of every tile in slot order (the order the one-device solve kernel uses) into the exchange
// buffer  [tile][KP][128] | G (KP x KP) , which the host all-reduces over the ranks (one collective per iteration,
// 8 (k n + k^2) bytes).  The buffer has the layout of a partial array with ONE slot per tile, so the H-step solve
// kernel reads it unchanged.  The W-step is local to a rank (its rows of W); only its three iteration scalars are
// summed over the ranks (second, 32-byte collective).
// (W'W)^-1 of a peer-memory session for ranks up to 32, by ONE extra block of the fold kernel: the ranks' side jobs
// published their W_q'W_q a hundred microseconds ago, so nothing waits here; the sum (rank order) and the
// Gauss-Jordan elimination (gauss_jordan_inverse, as in the side job of a one-device session: bit-identical G^-1)
// run on an idle SM next to the folding blocks and the rendezvous of their last block, off the critical path --
// instead of on side-job warps that share their issue slots with the streaming warps, where the chain
// publish -> wait -> remote loads -> sum -> invert stretched the stream kernel by 10-30 us.
__device__ __noinline__ void shard_gram_sum_invert(const HalfParams &p, int KP) {
    // 128 threads, the matrix (up to 32 x 32) in registers: thread t owns the entries (i = t / 4, j = 8 (t % 4) .. + 7).
    // Per pivot the owners put row p and column p into a double-buffered shared strip, ONE barrier, and every thread
    // updates its entries -- the arithmetic and the acceptance rule of gauss_jordan_inverse / als_shard_invert_kernel:
    //   a_ij - ((a_ip * ipiv) * a_pj) off the pivot cross, -a_ip * ipiv down column p, a_pj * ipiv along row p, ipiv at (p, p)
    // (~0.15 us per pivot; with the matrix in shared memory and two barriers per pivot this block was the tail of the
    // launch on one rank: 16 us at k = 16)
    __shared__ double rowb[2][32], colb[2][32];
    __shared__ unsigned long long nrm[2];
    const ShardPeers *sp = p.peers;
    const int t = threadIdx.x, k = p.k, par = p.epoch & 1, world = sp->world;
    if (t < world) shard_wait_flag((const int *)(sp->base[sp->rank] + SHARD_OFF_FLAGSG) + t, p.epoch, t, p.ctrl);
    if (t < 2) nrm[t] = 0ull;
    __syncthreads();
    if (*(volatile int *)&p.ctrl->stop) return;          // a peer never arrived (or the previous iteration has just converged)
    const int i = t >> 2, jq = t & 3, j0 = jq * 8;
    double v[8];
    double rs = 0.0;
#pragma unroll
    for (int u0 = 0; u0 < 8; u0 += 2) {
        double a[2] = {0.0, 0.0};
        for (int q0 = 0; q0 < world; q0 += 8) {          // the ranks' loads of two entries in flight, added in rank order
            double g[2][8];
#pragma unroll
            for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    g[uu][q] = (q0 + q < world && i < k && j0 + u0 + uu < k) ? ldcg(shard_gram(sp, q0 + q, par) + i * KP + j0 + u0 + uu) : 0.0;
#pragma unroll
            for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                for (int q = 0; q < 8; ++q) a[uu] += g[uu][q];
        }
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
            v[u0 + uu] = a[uu];
            if (i < KP && j0 + u0 + uu < KP) p.G[i * KP + j0 + u0 + uu] = a[uu];
            rs += fabs(a[uu]);
        }
    }
    // 1-norm of the symmetric matrix = largest absolute row sum (the 4 owners of a row are neighbouring lanes)
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    if (jq == 0 && i < k) atomicMax(&nrm[0], (unsigned long long)__double_as_longlong(rs));
    int sing = 0;
    for (int pv = 0; pv < k; ++pv) {
        const int b = pv & 1;
        if (i == pv) {
#pragma unroll
            for (int u = 0; u < 8; ++u) rowb[b][j0 + u] = v[u];
        }
        if (jq == (pv >> 3)) {
            double cv = v[0];
#pragma unroll
            for (int u = 1; u < 8; ++u) cv = ((pv & 7) == u) ? v[u] : cv;
            colb[b][i] = cv;
        }
        __syncthreads();
        double piv = rowb[b][pv];
        if (!(piv > 0.0) || !isfinite(piv)) { sing = 1; piv = 1.0; }
        const double ipiv = 1.0 / piv;
        const double f = colb[b][i] * ipiv;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            const double rj = rowb[b][j];
            if (i != pv) v[u] = (j != pv) ? v[u] - f * rj : -f;
            else v[u] = (j != pv) ? rj * ipiv : ipiv;
        }
    }
    rs = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        const bool in = i < k && j < k;
        if (i < KP && j < KP) p.Minv[i * KP + j] = in ? v[u] : 0.0;
        if (in) rs += fabs(v[u]);
    }
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    if (jq == 0 && i < k) atomicMax(&nrm[1], (unsigned long long)__double_as_longlong(rs));
    __syncthreads();
    if (t == 0) {
        const double an = __longlong_as_double((long long)nrm[0]), in = __longlong_as_double((long long)nrm[1]);
        if (!isfinite(in) || an * in * 2.220446049250313e-16 > 1.0) sing = 1;
        if (sing) {
            // (the close block of the same launch may find that the previous iteration converged: that takes precedence, as
            // in R, whichever block comes first -- it overwrites the reason, this one only sets it if there is none)
            atomicCAS(&p.ctrl->reason, 0, 3);
            __threadfence();
            p.ctrl->stop = 1;
        }
    }
}

__global__ void __launch_bounds__(TILE_UNITS) als_shard_fold_kernel(const HalfParams p, int KP, double *__restrict__ xchg, int invert_block,
                                                                    int close_block) {
    pdl_prologue(2);
    if (*(volatile int *)&p.ctrl->stop) return;
    if ((int)blockIdx.x == invert_block) {
        shard_gram_sum_invert(p, KP);
        return;
    }
    if ((int)blockIdx.x == close_block) {
        // The previous iteration is closed HERE, by one more block of this launch instead of a launch of its own (every
        // flag has long arrived).  If it converged, this launch has folded and exchanged W'A of an iteration R would not
        // have started -- on every rank alike, and the kernels after it see the decision and return.
        if (threadIdx.x < 32) shard_close_iteration(p, threadIdx.x);
        return;
    }
    const int nfold = (int)gridDim.x - (invert_block >= 0 ? 1 : 0) - (close_block >= 0 ? 1 : 0);
    const int tile = (int)blockIdx.x / KP, c = (int)blockIdx.x % KP, u = threadIdx.x;
    const long long tb = (long long)tile * p.SPT;
    int expected;
    if (p.U >= p.Gstream)
        expected = ALS_NG * (cta_of_unit(tb + p.SPT - 1, p.U, p.Gstream) - cta_of_unit(tb, p.U, p.Gstream) + 1);
    else
        expected = ALS_NG * p.SPT;
    double r = 0.0;
    if (c < p.k && (long long)tile * TILE_UNITS + u < p.nvalid) {
        const double *Pt = p.P + ((size_t)tile * p.maxslots * KP + c) * (size_t)TILE_UNITS + u;
        if (expected <= 12) {                                // all slots in flight at once, summed in slot order
            double v[12];
#pragma unroll
            for (int s = 0; s < 12; ++s) v[s] = (s < expected) ? ldcg(Pt + (size_t)s * (KP * TILE_UNITS)) : 0.0;
#pragma unroll
            for (int s = 0; s < 12; ++s) r += v[s];
        } else {
#pragma unroll 4
            for (int s = 0; s < expected; ++s) r += ldcg(Pt + (size_t)s * (KP * TILE_UNITS));
        }
    }
    const ShardPeers *sp = p.peers;
    const size_t xoff = ((size_t)tile * KP + c) * TILE_UNITS + u;
    if (sp && sp->push) {
        // slot `rank` of every rank's block (this rank's included): posted stores over NVLink, all in flight
        const int par = p.epoch & 1, me = sp->rank;
#pragma unroll
        for (int q = 0; q < SHARD_MAXP; ++q)
            if (q < sp->world) shard_xchg(sp, q, par, me)[xoff] = r;
    } else {
        xchg[xoff] = r;
    }
    if (!sp) {                                               // callback exchange: W_r'W_r rides in the same buffer
        if (blockIdx.x == 0)
            for (int e = u; e < KP * KP; e += TILE_UNITS) xchg[(size_t)p.T * KP * TILE_UNITS + e] = ldcg(&p.G[e]);
        return;
    }
    if (u == 0) atomicMin((unsigned long long *)(sp->base[sp->rank] + SHARD_OFF_STAMPS) + 4, globaltimer_ns());
    // peer mode.  The last block to finish publishes the buffer to every rank (itself included) and waits until all
    // ranks have published theirs: the H-step solve kernel, next in the stream, reads them over NVLink.
    __shared__ int is_last;
    __syncthreads();                                         // the block's stores happen-before thread 0's fence
    if (u == 0) {
        int *counter = (int *)(sp->base[sp->rank] + SHARD_OFF_COUNTER);
        if (sp->push) __threadfence_system();               // the block's stores went to the peers' memories
        else __threadfence();
        const int last = atomicAdd(counter, 1) == nfold - 1;
        if (last) *counter = 0;
        is_last = last;
    }
    __syncthreads();
    if (!is_last) return;
    shard_publish(sp, 0 /* flagsA */, p.epoch, u);
    unsigned long long *stamp = (unsigned long long *)(sp->base[sp->rank] + SHARD_OFF_STAMPS);
    if (u == 0) stamp[0] = globaltimer_ns();
    if (u < sp->world) shard_wait_flag((const int *)sp->base[sp->rank] + u, p.epoch, u, p.ctrl);
    __syncthreads();
    if (u == 0) stamp[1] = globaltimer_ns();
}

// callback exchange, after the all-reduce: G = W'W of the whole matrix -> p.G, its inverse -> p.Minv (R's solve()
// failure: reason 3).  (Peer-memory exchange: the Gram side job of the stream kernel does this, see gram_side_job.)
// One block of 1024 threads, the whole matrix in registers: thread t owns the entries (i = t / 16, j = 4 (t % 16) .. + 3).
// Per pivot the owners publish row p and column p to a double-buffered shared strip, ONE barrier, and every thread
// updates its four entries -- the same Gauss-Jordan arithmetic and acceptance rule as gauss_jordan_inverse (nnls.cuh):
//   a_ij - ((a_ip * ipiv) * a_pj) off the pivot cross, -a_ip * ipiv down column p, a_pj * ipiv along row p, ipiv at (p, p).
// ~0.2 us per pivot instead of ~1 us with the matrix in shared memory and two barriers per pivot.
__global__ void __launch_bounds__(1024) als_shard_invert_kernel(const HalfParams p, int KP, const double *Gsum, int reason) {
    __shared__ double rowb[2][64], colb[2][64];
    __shared__ unsigned long long nrm[2];
    pdl_prologue();
    if (*(volatile int *)&p.ctrl->stop) return;
    const int k = p.k, t = threadIdx.x;
    const int i = t >> 4, jq = t & 15, j0 = jq * 4;
    double v[4];
    double rs = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        const bool in = i < k && j < k;
        v[u] = in ? Gsum[i * KP + j] : 0.0;
        if (i < KP && j < KP) p.G[i * KP + j] = v[u];
        rs += fabs(v[u]);
    }
    if (t < 2) nrm[t] = 0ull;
    __syncthreads();
    // 1-norm of the symmetric matrix = largest absolute row sum (the 16 owners of a row are half a warp)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (jq == 0 && i < k) atomicMax(&nrm[0], (unsigned long long)__double_as_longlong(rs));
    int sing = 0;
    for (int pv = 0; pv < k; ++pv) {
        const int b = pv & 1;
        if (i == pv) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rowb[b][j0 + u] = v[u];
        }
        if (jq == (pv >> 2)) {
            const int u = pv & 3;
            colb[b][i] = (u == 0) ? v[0] : (u == 1) ? v[1] : (u == 2) ? v[2] : v[3];
        }
        __syncthreads();
        double piv = rowb[b][pv];
        if (!(piv > 0.0) || !isfinite(piv)) { sing = 1; piv = 1.0; }
        const double ipiv = 1.0 / piv;
        const double f = colb[b][i] * ipiv;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            const double rj = rowb[b][j];
            if (i != pv) v[u] = (j != pv) ? v[u] - f * rj : -f;
            else v[u] = (j != pv) ? rj * ipiv : ipiv;
        }
    }
    rs = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        const bool in = i < k && j < k;
        if (i < KP && j < KP) p.Minv[i * KP + j] = in ? v[u] : 0.0;
        if (in) rs += fabs(v[u]);
    }
    // 1-norm of the inverse from its absolute ROW sums (the computed inverse is symmetric up to rounding, and the norm only
    // feeds the rcond threshold an * in * eps > 1)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (jq == 0 && i < k) atomicMax(&nrm[1], (unsigned long long)__double_as_longlong(rs));
    __syncthreads();
    if (t == 0) {
        const double an = __longlong_as_double((long long)nrm[0]), in = __longlong_as_double((long long)nrm[1]);
        if (!isfinite(in) || an * in * 2.220446049250313e-16 > 1.0) sing = 1;
        if (sing) {
            p.ctrl->reason = reason;
            __threadfence();
            p.ctrl->stop = 1;
        }
    }
}

// Peer-memory sessions, between the H-step's stream kernel and its fold kernel: closes the previous iteration
// (als_shard_close_kernel's job, warp 0) and forms (W'W)^-1 from the ranks' W_q'W_q, which their side jobs published
// a hundred microseconds ago: no rank waits here, the rendezvous is off the critical path, and the serial chain
// wait -> remote loads -> sum -> Gauss-Jordan runs on an idle SM in ~5 us instead of on side-job warps that share
// their issue slots with the streaming warps (where it stretched the stream kernel by 10-25 us).  The sum is in rank
// order and the elimination is als_shard_invert_kernel's: every rank gets bit-identical G and G^-1.
__global__ void __launch_bounds__(1024) als_shard_close_invert_kernel(const HalfParams p, int KP) {
    __shared__ double rowb[2][64], colb[2][64];
    __shared__ unsigned long long nrm[2];
    pdl_prologue();
    AlsCtrl *c = p.ctrl;
    const ShardPeers *sp = p.peers;
    const int k = p.k, t = threadIdx.x;
    // a batch that stopped earlier (converged, zero row, ...): the stream kernel before this one returned at once and
    // published nothing -- identically on every rank
    if (*(volatile int *)&c->stop) return;
    if (t < 32) {
        shard_close_iteration(p, t);                     // the iteration whose W-step scalars were written last
    } else if (t < 32 + sp->world) {
        // ---- every rank's W_q'W_q of this epoch has been published ----
        const int q = t - 32;
        shard_wait_flag((const int *)(sp->base[sp->rank] + SHARD_OFF_FLAGSG) + q, p.epoch, q, c);
    }
    if (t < 2) nrm[t] = 0ull;
    __syncthreads();
    if (*(volatile int *)&c->stop) return;               // converged just now, failed earlier, or a peer never arrived
    const int i = t >> 4, jq = t & 15, j0 = jq * 4;
    double v[4];
    double rs = 0.0;
    {
        const int par = p.epoch & 1, world = sp->world;
        // the ranks' loads of two entries in flight at once (64 registers per thread at 1024 threads), added in rank order
#pragma unroll
        for (int u0 = 0; u0 < 4; u0 += 2) {
            double a[2] = {0.0, 0.0};
            for (int q0 = 0; q0 < world; q0 += 8) {
                double g[2][8];
#pragma unroll
                for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        g[uu][q] = (q0 + q < world && i < k && j0 + u0 + uu < k) ? ldcg(shard_gram(sp, q0 + q, par) + i * KP + j0 + u0 + uu) : 0.0;
#pragma unroll
                for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                    for (int q = 0; q < 8; ++q) a[uu] += g[uu][q];
            }
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
                v[u0 + uu] = a[uu];
                if (i < KP && j0 + u0 + uu < KP) p.G[i * KP + j0 + u0 + uu] = a[uu];
                rs += fabs(a[uu]);
            }
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (jq == 0 && i < k) atomicMax(&nrm[0], (unsigned long long)__double_as_longlong(rs));
    int sing = 0;
    for (int pv = 0; pv < k; ++pv) {
        const int b = pv & 1;
        if (i == pv) {
#pragma unroll
            for (int u = 0; u < 4; ++u) rowb[b][j0 + u] = v[u];
        }
        if (jq == (pv >> 2)) {
            const int u = pv & 3;
            colb[b][i] = (u == 0) ? v[0] : (u == 1) ? v[1] : (u == 2) ? v[2] : v[3];
        }
        __syncthreads();
        double piv = rowb[b][pv];
        if (!(piv > 0.0) || !isfinite(piv)) { sing = 1; piv = 1.0; }
        const double ipiv = 1.0 / piv;
        const double f = colb[b][i] * ipiv;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            const double rj = rowb[b][j];
            if (i != pv) v[u] = (j != pv) ? v[u] - f * rj : -f;
            else v[u] = (j != pv) ? rj * ipiv : ipiv;
        }
    }
    rs = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        const bool in = i < k && j < k;
        if (i < KP && j < KP) p.Minv[i * KP + j] = in ? v[u] : 0.0;
        if (in) rs += fabs(v[u]);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (jq == 0 && i < k) atomicMax(&nrm[1], (unsigned long long)__double_as_longlong(rs));
    __syncthreads();
    if (t == 0) {
        const double an = __longlong_as_double((long long)nrm[0]), in = __longlong_as_double((long long)nrm[1]);
        if (!isfinite(in) || an * in * 2.220446049250313e-16 > 1.0) sing = 1;
        if (sing) {
            c->reason = 3;
            __threadfence();
            c->stop = 1;
        }
    }
}

__global__ void __launch_bounds__(32) als_shard_finish_kernel(const HalfParams p) {
    pdl_prologue();
    AlsCtrl *c = p.ctrl;
    if (*(volatile int *)&c->stop) return;
    if (threadIdx.x == 0) finish_iteration(c, p.trace, p.xscal[0], p.xscal[1], p.xscal[2]);
}

// r = F a  style helper for mfb_nnls: R[c][j] = sum_i C[i][c] B[i][j]  (small, test entry point)
__global__ void __launch_bounds__(256) small_crossprod_kernel(const double *__restrict__ C, const double *__restrict__ B,
                                                              int64_t r, int k, int64_t pcols, double *__restrict__ R) {
    // one block per column j of B; thread c2 < k ... generic slow path, test-only sizes
    const int64_t j = blockIdx.x;
    __shared__ double red[256];
    for (int c = 0; c < k; ++c) {
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < r; i += 256) s += C[c * r + i] * B[j * r + i];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) R[j * k + c] = red[0];
        __syncthreads();
    }
}

// mfb_nnls: the product solver (nnls_warp) on given normal equations, 8 columns per CTA
template <int KB>
__global__ void __launch_bounds__(128) nnls_columns_kernel(const double *__restrict__ G, const double *__restrict__ Minv,
                                                           const double *__restrict__ R, int k, int64_t pcols,
                                                           double *__restrict__ K, unsigned long long *draws) {
    constexpr int KP = KB * 8, SU = 8, XS = SU + 1;
    extern __shared__ double nn_sh[];
    double *sG = nn_sh, *sM = sG + KP * KP, *xs = sM + KP * KP, *scr = xs + KP * XS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < KP * KP; e += 128) {
        const int c1 = e / KP, c2 = e - c1 * KP;
        sG[e] = (c1 < k && c2 < k) ? G[c1 * KP + c2] : 0.0;
        sM[e] = (c1 < k && c2 < k) ? Minv[c1 * KP + c2] : 0.0;
    }
    const int64_t j0 = blockIdx.x * (int64_t)SU;
    const int nvalid = (int)((pcols - j0 < SU) ? pcols - j0 : SU);
    for (int e = tid; e < SU * KP; e += 128) {
        const int u = e % SU, c = e / SU;
        xs[c * XS + u] = (c < k && u < nvalid) ? R[(j0 + u) * k + c] : 0.0;
    }
    __syncthreads();
    if constexpr (KP <= 32) nnls_lanes<KP, XS>(sG, sM, k, xs, warp * 2, 2, nvalid, lane, scr + warp * NnlsCfg<KP>::SCR, draws);
    else nnls_warp<KP, XS>(sG, sM, k, xs, scr + warp * NnlsCfg<KP>::SCR, warp * 2, 2, nvalid, lane);
    __syncthreads();
    for (int e = tid; e < SU * KP; e += 128) {
        const int u = e % SU, c = e / SU;
        if (c < k && u < nvalid) K[(j0 + u) * k + c] = xs[c * XS + u];
    }
}

template <int KB>
static int launch_nnls_columns(const double *G, const double *Minv, const double *R, int k, int64_t pcols, double *K,
                               cudaStream_t st, unsigned long long *draws) {
    constexpr int KP = KB * 8;
    const size_t smem = (size_t)(2 * KP * KP + KP * 9 + 4 * NnlsCfg<KP>::SCR) * sizeof(double);
    MFB_CUDA(cudaFuncSetAttribute(nnls_columns_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nnls_columns_kernel<KB><<<(unsigned)((pcols + 7) / 8), 128, smem, st>>>(G, Minv, R, k, pcols, K, draws);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// snmf: the convergence test of NMF::.nmf_snmf (every fifth iteration; R/bicluster.R:363 delegates to it)
// ---------------------------------------------------------------------------
// One launch per factor X (H after the stream kernel of the NEXT H-step has run ahead, W after its own W-step): for the
// units [128 b, 128 b + 128) of block b
//   res = pmin(x, G x - r)   G = the regularised Gram matrix the half-step's solve kernel uses (W'W + beta 11' or
//                            HH' + eta^2 I), r = W'A / A H' = the stream-K partials summed in slot order
// and part[b][0..4] = sum |res|, #{|res| > 0}, #{arg-max membership != previous test's}, sum x^2, sum (sum_c x_c)^2.
// The memberships are max.col(ties.method = "first"), 1-based like R's (the previous ones start as 0).
#define SNMF_CHECK_Y 4            // rank slices per unit: blockDim = (128 units, 4)
__global__ void __launch_bounds__(TILE_UNITS * SNMF_CHECK_Y) snmf_check_kernel(const HalfParams p, int KP, int with_grad,
                                                                               const int *__restrict__ idx_old, int *__restrict__ idx_new,
                                                                               double *__restrict__ part) {
    // thread (u, y) evaluates the entries c = y, y + 4, ... of unit u; the unit's factor values are staged in shared
    // memory (dynamic: KP x 128 doubles after the KP x KP Gram matrix) so that every slice sees all of x
    extern __shared__ double chk_sh[];
    double *sG = chk_sh, *sx = chk_sh + KP * KP;
    __shared__ double sred[TILE_UNITS * SNMF_CHECK_Y / 32][5];
    const int u = threadIdx.x, y = threadIdx.y, tid = y * TILE_UNITS + u, tile = blockIdx.x, k = p.k;
    const long long unit = (long long)tile * TILE_UNITS + u;
    const bool valid = unit < p.nvalid;
    for (int e = tid; e < KP * KP; e += TILE_UNITS * SNMF_CHECK_Y) sG[e] = with_grad ? p.G[e] : 0.0;
    for (int c = y; c < k; c += SNMF_CHECK_Y) sx[c * TILE_UNITS + u] = valid ? p.Fout[(size_t)c * p.ldout + unit] : 0.0;
    __syncthreads();
    double sabs = 0.0, cnt = 0.0, changed = 0.0, sx2 = 0.0, cs2 = 0.0;
    if (valid && y == 0) {
        // memberships (first maximum), sum x^2, (sum_c x_c)^2: one thread per unit, c in order
        int best = 0;
        double bx = sx[u], cs = 0.0;
        for (int c = 0; c < k; ++c) {
            const double x = sx[c * TILE_UNITS + u];
            sx2 += x * x;
            cs += x;
            if (x > bx) { bx = x; best = c; }
        }
        cs2 = cs * cs;
        idx_new[unit] = best + 1;
        changed = (idx_old[unit] != best + 1) ? 1.0 : 0.0;
    }
    if (valid && with_grad) {
        const int expected = expected_slots(p, tile);
        for (int c = y; c < k; c += SNMF_CHECK_Y) {
            const double *Pt = p.P + ((size_t)tile * p.maxslots * KP + c) * (size_t)TILE_UNITS + u;
            double r = 0.0;
            for (int sl = 0; sl < expected; ++sl) r += ldcg(Pt + (size_t)sl * (KP * TILE_UNITS));
            double g = 0.0;
            for (int j = 0; j < k; ++j) g += sG[j * KP + c] * sx[j * TILE_UNITS + u];   // G symmetric: broadcast reads
            const double res = fabs(fmin(sx[c * TILE_UNITS + u], g - r));
            sabs += res;
            if (res > 0.0) cnt += 1.0;
        }
    }
    double v[5] = {sabs, cnt, changed, sx2, cs2};
#pragma unroll
    for (int q = 0; q < 5; ++q) v[q] = warp_sum(v[q]);
    if ((tid & 31) == 0)
        for (int q = 0; q < 5; ++q) sred[tid >> 5][q] = v[q];
    __syncthreads();
    if (tid < 5) {
        double a = 0.0;
        for (int w = 0; w < TILE_UNITS * SNMF_CHECK_Y / 32; ++w) a += sred[w][tid];      // warp order: deterministic
        part[(size_t)tile * 8 + tid] = a;
    }
}

struct mfb_als {
    mfb_context *ctx;
    mfb_matrix *A;
    int k, KB, KP;
    int64_t m, n, ldw, ldh;
    double *Wb[2], *Wr, *Hb[2];          // factor buffers [KP][ld]
    double *G_W, *Minv_W, *G_H, *Minv_H; // Grams (+ inverses)
    double *P_h, *P_w, *spart, *spart2, *gram_part, *gram_part2;
    int *counters, *gram_counter;
    int nsolve_h, nsolve_w;       // solve-kernel grids
    bool attr_set = false;
    bool f32 = false;             // A is resident in single precision
    bool tf32 = false;            // ... and the products run on the TF32 tensor cores (als_tf32.cuh)
    int KPN = 0;                  // rank padded to the MMA N granularity (16)
    float *Whi[3] = {nullptr, nullptr, nullptr}, *Wlo[3] = {nullptr, nullptr, nullptr};   // hi / lo FP32 copies of Wb[0], Wb[1], Wr
    float *Hhi[2] = {nullptr, nullptr}, *Hlo[2] = {nullptr, nullptr};
    CUtensorMap tmf_Whi[3], tmf_Wlo[3], tmf_Hhi[2], tmf_Hlo[2];
    size_t smem_h32 = 0, smem_w32 = 0;
    int nstage_h32 = 0, nstage_w32 = 0;
    std::vector<cudaEvent_t> *prof = nullptr;   // mfb_als_profile: one event after every kernel
    AlsCtrl *ctrl;
    double *trace_dev;
    int maxslots_h, maxslots_w, T_h, T_w, SPT_h, SPT_w;
    int nstage_h, nstage_w;
    size_t smem_h, smem_w;
    CUtensorMap tm_h, tm_w;
    CUtensorMap tm_W[3], tm_H[2];   // factor slabs: Wb[0], Wb[1], Wr / Hb[0], Hb[1]
    int cur;              // index of the buffer holding the last accepted W / H
    bool restart_pending; // next H-step reads Wr instead of Wb[cur]
    int iters_total;      // R's `i`
    int max_trace;
    std::vector<double> trace;   // host copy: iters_total x 3
    double resid;
    std::vector<void *> allocs;
    // row-sharded mode (mfb_als_begin_sharded): this session holds one row block of the matrix
    bool shard = false;
    mfb_allreduce_fn allreduce = nullptr;
    void *allreduce_user = nullptr;
    int64_t m_global = 0;
    double *xchg = nullptr, *xscal = nullptr;   // [T_h][KP][128] + KP*KP exchange buffer; [4] scalars
    // ... over peer memory (mfb_als_shard_export / mfb_als_shard_connect)
    unsigned char *shared = nullptr;            // this rank's shared block (cudaMalloc, cudaIpc-exported)
    size_t xchg_cnt = 0;
    int shard_slots = 1;                        // exchange buffers per parity in the shared block (push exchange up to this many ranks)
    ShardPeers peers_host;
    ShardPeers *peers_dev = nullptr;
    std::vector<void *> peer_maps;              // cudaIpcOpenMemHandle mappings to close
    int epoch = 0;
    // snmf sessions (mfb_snmf_begin): NMF::nmf(method = "snmf/r") on the same kernels with regularised Gram matrices
    bool snmf = false;
    double beta = 0.0, eta = 0.0;
    bool prestreamed = false;     // the stream kernel of the next H-step has already run (for the convergence test)
    int inc = 0;                  // consecutive tests with stable arg-max memberships
    double erravg1 = -1.0;        // first mean KKT residual since the last (re)start (< 0: not set)
    int *idxW[2] = {nullptr, nullptr}, *idxH[2] = {nullptr, nullptr};   // arg-max memberships: [idx_cur] = previous test
    int idx_cur = 0;
    double *chk_part = nullptr;   // [blocks of both check launches][8]
    double objective = NAN;
};

static int max_slots(long long T, long long SPT, int G) {
    const long long U = T * SPT;
    if (U < G) return (int)SPT;
    int mx = 1;
    for (long long t = 0; t < T; ++t) {
        long long a = ((t * SPT + 1) * G + U - 1) / U - 1;
        long long b = (((t + 1) * SPT) * G + U - 1) / U - 1;
        if ((int)(b - a + 1) > mx) mx = (int)(b - a + 1);
    }
    return mx;
}

static int encode_map(CUtensorMap *tm, void *base, int64_t rows, int64_t cols, int64_t lda, int box_rows, int box_cols,
                      bool f32 = false) {
    mfb_encode_tiled_fn enc = mfb_get_encode_tiled();
    if (!enc) {
        mfb_set_error("cuTensorMapEncodeTiled unavailable");
        return MFB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)lda * (f32 ? sizeof(float) : sizeof(double))};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        mfb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return MFB_ERR_CUDA;
    }
    return MFB_OK;
}

static int encode_map_rowgroups(CUtensorMap *tm, void *base, int64_t cols, int64_t lda, bool f32 = false) {
    // 3-D view of the column-major matrix: {16 rows, cols, lda/16 row groups}, strides {lda*8, 128} bytes;
    // box {16, 32, 8} = a 128-row x 32-column tile that lands as [row group][column][16 rows]
    // (FP32: {32 rows, cols, lda/32 groups}, box {32, 32, 4}: the same 128-byte rows).
    mfb_encode_tiled_fn enc = mfb_get_encode_tiled();
    if (!enc) { mfb_set_error("cuTensorMapEncodeTiled unavailable"); return MFB_ERR_CUDA; }
    const int per = f32 ? 32 : 16;
    cuuint64_t dims[3] = {(cuuint64_t)per, (cuuint64_t)cols, (cuuint64_t)(lda / per)};
    cuuint64_t strides[2] = {(cuuint64_t)lda * (f32 ? sizeof(float) : sizeof(double)), 128};
    cuuint32_t box[3] = {(cuuint32_t)per, 32, (cuuint32_t)(128 / per)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { mfb_set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r); return MFB_ERR_CUDA; }
    return MFB_OK;
}

#define MFB_TRY(expr) do { int _r = (expr); if (_r != MFB_OK) return _r; } while (0)

static int prof_mark(mfb_als *s) {
    if (!s->prof) return MFB_OK;
    cudaEvent_t e;
    MFB_CUDA(cudaEventCreate(&e));
    MFB_CUDA(cudaEventRecord(e, s->ctx->stream));
    s->prof->push_back(e);
    return MFB_OK;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl,
                              Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

static bool als_peer_mode(const mfb_als *s) {
    return s->shard && s->peers_host.world > 0 && s->peer_maps.size() + 1 == (size_t)s->peers_host.world;
}
static bool als_use_pdl(const mfb_als *s) {
    // events between launches serialise anyway; with the callback exchange collectives sit between the kernels
    return (s->prof == nullptr) && (!s->shard || als_peer_mode(s)) && !getenv("MFB_NO_PDL");
}

// TF32 path (als_tf32.cuh): tensor maps of the fixed factor's hi / lo FP32 copies, and where the factor produced by
// this half-step is to be split into
struct Tf32Extra {
    const CUtensorMap *bhi = nullptr, *blo = nullptr;
    float *out_hi = nullptr, *out_lo = nullptr;
};

template <int KB, bool F32>
static int launch_half_t(mfb_als *s, bool wstep, const HalfParams &hp, const CUtensorMap &tmF, int which, const Tf32Extra &x) {
    // which: 0 = stream + solve, 1 = stream kernel only, 2 = solve kernel only (row-sharded H-step)
    cudaStream_t st = s->ctx->stream;
    const int G = hp.Gstream;
    const int nsolve = wstep ? s->nsolve_w : s->nsolve_h;
    const bool pdl = als_use_pdl(s);
    constexpr bool TF32_OK = F32 && (KB <= 2);
    const bool tf32 = TF32_OK && s->tf32 && x.bhi != nullptr;
    if (!s->attr_set) {
        MFB_CUDA(cudaFuncSetAttribute(als_stream_kernel<KB, true, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_w));
        MFB_CUDA(cudaFuncSetAttribute(als_stream_kernel<KB, false, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_h));
        MFB_CUDA(cudaFuncSetAttribute(als_solve_kernel<KB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SolveCfg<KB>::SMEM));
        MFB_CUDA(cudaFuncSetAttribute(als_solve_kernel<KB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SolveCfg<KB>::SMEM));
        if constexpr (TF32_OK) {
            if (s->tf32) {
                MFB_CUDA(cudaFuncSetAttribute(als_stream_tf32_kernel<KB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_w32));
                MFB_CUDA(cudaFuncSetAttribute(als_stream_tf32_kernel<KB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_h32));
            }
        }
        s->attr_set = true;
    }
    auto stream_launch = [&](bool w) -> int {
        if constexpr (TF32_OK) {
            if (tf32) {
                if (w) MFB_CUDA(launch_pdl(als_stream_tf32_kernel<KB, true>, G, TF32_THREADS, s->smem_w32, st, pdl, s->tm_w, *x.bhi, *x.blo, hp, s->nstage_w32));
                else MFB_CUDA(launch_pdl(als_stream_tf32_kernel<KB, false>, G, TF32_THREADS, s->smem_h32, st, pdl, s->tm_h, *x.bhi, *x.blo, hp, s->nstage_h32));
                return MFB_OK;
            }
        }
        if (w) MFB_CUDA(launch_pdl(als_stream_kernel<KB, true, F32>, G, ALS_THREADS, s->smem_w, st, pdl, s->tm_w, tmF, hp, s->nstage_w));
        else MFB_CUDA(launch_pdl(als_stream_kernel<KB, false, F32>, G, ALS_THREADS, s->smem_h, st, pdl, s->tm_h, tmF, hp, s->nstage_h));
        return MFB_OK;
    };
    auto split_launch = [&]() -> int {      // hi / lo FP32 copies of the factor the solve kernel has just written
        if (!(s->tf32 && x.out_hi)) return MFB_OK;
        const int64_t total = (int64_t)s->KP * hp.ldout;
        int nb = (int)((total + 255) / 256);
        if (nb > 2 * s->ctx->sm_count) nb = 2 * s->ctx->sm_count;
        MFB_CUDA(launch_pdl(als_split_factor_kernel, nb, 256, 0, st, pdl, (const double *)hp.Fout, hp.ldout, s->KP, x.out_hi,
                            x.out_lo, (const AlsCtrl *)hp.ctrl));
        return MFB_OK;
    };
    auto invert_launch = [&](int reason) -> int {   // (timed with the solve kernel that follows)
        if (!hp.invert_later) return MFB_OK;
        MFB_CUDA(launch_pdl(als_shard_invert_kernel, 1, 1024, 0, st, pdl, hp, s->KP,
                            (const double *)hp.G, reason));
        return MFB_OK;
    };
    if (wstep) {
        MFB_TRY(stream_launch(true));
        MFB_TRY(prof_mark(s));
        MFB_TRY(invert_launch(4));
        MFB_CUDA(launch_pdl(als_solve_kernel<KB, true>, nsolve, SOLVE_THREADS, SolveCfg<KB>::SMEM, st, pdl, hp));
        MFB_TRY(split_launch());
        MFB_TRY(prof_mark(s));
    } else {
        if (which != 2) {
            MFB_TRY(stream_launch(false));
            MFB_TRY(prof_mark(s));
            MFB_TRY(invert_launch(3));
        }
        if (which != 1) {
            MFB_CUDA(launch_pdl(als_solve_kernel<KB, false>, nsolve, SOLVE_THREADS, SolveCfg<KB>::SMEM, st, pdl, hp));
            MFB_TRY(split_launch());
            MFB_TRY(prof_mark(s));
        }
    }
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

template <int KB>
static int launch_half(mfb_als *s, bool wstep, const HalfParams &hp, const CUtensorMap &tmF, int which, const Tf32Extra &x) {
    return s->f32 ? launch_half_t<KB, true>(s, wstep, hp, tmF, which, x) : launch_half_t<KB, false>(s, wstep, hp, tmF, which, x);
}

template <int KB>
static int solve_units(void) { return SolveCfg<KB>::SU; }

static int solve_units_dispatch(int KB) {
    switch (KB) {
        case 1: return solve_units<1>();
        case 2: return solve_units<2>();
        case 3: return solve_units<3>();
        case 4: return solve_units<4>();
        case 6: return solve_units<6>();
        case 8: return solve_units<8>();
    }
    return 0;
}

static int check_mbar_abort(const char *where) {
    int flag = 0, rec[8];
    if (cudaMemcpyFromSymbol(&flag, g_mbar_abort, sizeof(int)) != cudaSuccess) return MFB_OK;
    if (!flag) return MFB_OK;
    cudaMemcpyFromSymbol(rec, g_mbar_record, sizeof(rec));
    int zero = 0;
    cudaMemcpyToSymbol(g_mbar_abort, &zero, sizeof(int));
    mfb_set_error("%s: mbarrier wait timed out (kind %d [1=producer/empty, 2=consumer/full], block %d, thread %d, unit %d, stage %d, parity %d)",
                  where, rec[0], rec[1], rec[2], rec[3], rec[4], rec[5]);
    return MFB_ERR_CUDA;
}

static int launch_half_dispatch(mfb_als *s, bool wstep, const HalfParams &hp, const CUtensorMap &tmF, int which = 0,
                                const Tf32Extra &x = Tf32Extra()) {
    switch (s->KB) {
        case 1: return launch_half<1>(s, wstep, hp, tmF, which, x);
        case 2: return launch_half<2>(s, wstep, hp, tmF, which, x);
        case 3: return launch_half<3>(s, wstep, hp, tmF, which, x);
        case 4: return launch_half<4>(s, wstep, hp, tmF, which, x);
        case 6: return launch_half<6>(s, wstep, hp, tmF, which, x);
        case 8: return launch_half<8>(s, wstep, hp, tmF, which, x);
    }
    mfb_set_error("unsupported rank block %d", s->KB);
    return MFB_ERR_UNSUPPORTED;
}

template <typename T>
static int dev_alloc(mfb_als *s, T **p, size_t count, bool zero = true) {
    if (mfb_pool_alloc(s->ctx, (void **)p, count * sizeof(T)) != MFB_OK) return MFB_ERR_ALLOC;
    s->allocs.push_back((void *)*p);
    if (zero) MFB_CUDA(cudaMemsetAsync(*p, 0, count * sizeof(T), s->ctx->stream));
    return MFB_OK;
}

extern "C" {

void mfb_als_end(mfb_als *s) {
    if (!s) return;
    if (mfb_context_alive(s->ctx)) {
        cudaSetDevice(s->ctx->device);
        cudaStreamSynchronize(s->ctx->stream);
        for (void *p : s->allocs) mfb_pool_free(s->ctx, p);
        for (void *p : s->peer_maps) cudaIpcCloseMemHandle(p);
        if (s->shared) cudaFree(s->shared);
    }
    delete s;
}

static int als_upload_H(mfb_als *s, const double *H_host, double *dst) {
    // host H is k x n column-major; device layout is [c][j] (j contiguous, ld = ldh)
    std::vector<double> tmp((size_t)s->k * s->n);
    for (int64_t j = 0; j < s->n; ++j)
        for (int c = 0; c < s->k; ++c) tmp[(size_t)c * s->n + j] = H_host[(size_t)j * s->k + c];
    MFB_CUDA(cudaMemcpy2DAsync(dst, s->ldh * sizeof(double), tmp.data(), s->n * sizeof(double), s->n * sizeof(double),
                               s->k, cudaMemcpyHostToDevice, s->ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return MFB_OK;
}

static int als_begin_impl(mfb_context *ctx, mfb_matrix *A, int64_t m_global, int k, const double *W0, const double *Hold0,
                         double resid_old0, mfb_allreduce_fn allreduce, void *user, mfb_als **out) {
    MFB_ARG(ctx && A && W0 && Hold0 && out, "NULL argument");
    MFB_ARG(k >= 1 && k <= 64, "k must be in 1..64");
    const bool shard = allreduce != nullptr;
    MFB_ARG(k <= (shard ? m_global : A->m) && k <= A->n, "k exceeds a matrix dimension");
    MFB_ARG(!shard || (A->m >= 1 && A->m <= m_global), "row block larger than the matrix");
    MFB_CUDA(cudaSetDevice(ctx->device));
    double stats[5];
    MFB_TRY(mfb_matrix_stats(A, stats));
    if (!shard) {      // (sharded: the ranks agree on these through the first collective, below)
        if (stats[3] > 0) {
            mfb_set_error("als_nmf: matrix has NA/NaN entries");
            return MFB_ERR_ARG;
        }
        if (stats[0] < 0) {
            mfb_set_error("Negative values are not allowed");
            return MFB_ERR_NEGATIVE;
        }
    }
    mfb_als *s = new mfb_als();
    s->shard = shard;
    s->allreduce = allreduce;
    s->allreduce_user = user;
    s->m_global = shard ? m_global : A->m;
    s->ctx = ctx;
    s->A = A;
    s->k = k;
    int KB = (k + 7) / 8;
    if (KB == 5) KB = 6;
    if (KB == 7) KB = 8;
    s->KB = KB;
    s->KP = KB * 8;
    s->m = A->m;
    s->n = A->n;
    s->ldw = round_up(A->m, 128);
    s->ldh = round_up(A->n, 128);
    s->cur = 0;
    s->restart_pending = false;
    s->iters_total = 0;
    s->resid = resid_old0;
    const int G = ctx->sm_count;
    s->f32 = A->elem_bytes == 4;
    s->T_h = (int)((s->n + 127) / 128);
    s->SPT_h = (int)((s->m + (s->f32 ? 31 : 15)) / (s->f32 ? 32 : 16));
    s->T_w = (int)((s->m + 127) / 128);
    s->SPT_w = (int)((s->n + 31) / 32);
    s->maxslots_h = ALS_NG * max_slots(s->T_h, s->SPT_h, G);   // x NG consumer groups
    s->maxslots_w = ALS_NG * max_slots(s->T_w, s->SPT_w, G);
    const size_t KP = s->KP;
    const size_t side = (KP * KP + 2) * sizeof(double) + 16;            // Gram side job scratch
    const size_t budget = 227 * 1024 - 1024 /*align*/ - 2 * 16 * sizeof(uint64_t) - side;
    const size_t stage_h = 16384 + (s->f32 ? 2 : 1) * KP * 128, stage_w = (s->f32 ? 16384 : 32768) + 2 * KP * 128;
    s->nstage_h = (int)(budget / stage_h); if (s->nstage_h > 12) s->nstage_h = 12;
    s->nstage_w = (int)(budget / stage_w); if (s->nstage_w > (s->f32 ? 12 : 6)) s->nstage_w = s->f32 ? 12 : 6;
    // The ring depth must be a multiple of the number of consumer groups (2): a stage is then always drained by
    // the same group, which therefore waits on EVERY phase of that stage's full-barrier.  With an odd depth a
    // group skips every other phase and the parity wait aliases (returns on a phase two fills old).
    s->nstage_h -= s->nstage_h % ALS_NG;
    s->nstage_w -= s->nstage_w % ALS_NG;
    if (s->nstage_h < 2 || s->nstage_w < 2) {
        mfb_als_end(s);
        mfb_set_error("rank %d leaves no shared memory for the streaming ring", k);
        return MFB_ERR_UNSUPPORTED;
    }
    s->smem_h = 1024 + (size_t)s->nstage_h * stage_h + 2 * s->nstage_h * sizeof(uint64_t) + side;
    s->smem_w = 1024 + (size_t)s->nstage_w * stage_w + 2 * s->nstage_w * sizeof(uint64_t) + side;
    // FP32 matrix: the products run on the TF32 tensor cores (als_tf32.cuh) for ranks up to 16; MFB_F32_DMMA=1 keeps
    // the FP32-storage / FP64-tensor-core kernels (development aid for A/B runs)
    s->KPN = (int)round_up(s->KP, 16);
    s->tf32 = s->f32 && s->KB <= 2 && !getenv("MFB_F32_DMMA");
    if (s->tf32) {
        const int comp = getenv("MFB_TF32_NOCOMP") ? 0 : 1;
        cudaMemcpyToSymbol(g_tf32_comp, &comp, sizeof(int));
        const size_t stage32 = 16384 + 2 * (size_t)s->KPN * 128;
        const size_t bars32 = 16 * sizeof(uint64_t) + 16;
        int ns = (int)((227 * 1024 - 1024 - 2 * 16 * sizeof(uint64_t) - bars32 - side) / stage32);
        if (ns > 10) ns = 10;
        ns -= ns % 2;
        s->nstage_h32 = s->nstage_w32 = ns;
        s->smem_h32 = s->smem_w32 = 1024 + (size_t)ns * stage32 + 2 * ns * sizeof(uint64_t) + bars32 + side;
    }
    const int SU = solve_units_dispatch(s->KB);
    s->nsolve_h = (int)((s->n + SU - 1) / SU);
    s->nsolve_w = (int)((s->m + SU - 1) / SU);

    int rc;
#define ALLOC(ptr, cnt) if ((rc = dev_alloc(s, &(ptr), (cnt))) != MFB_OK) { mfb_als_end(s); return rc; }
    ALLOC(s->Wb[0], KP * s->ldw) ALLOC(s->Wb[1], KP * s->ldw) ALLOC(s->Wr, KP * s->ldw)
    ALLOC(s->Hb[0], KP * s->ldh) ALLOC(s->Hb[1], KP * s->ldh)
    if (s->tf32) {
        for (int i = 0; i < 3; ++i) { ALLOC(s->Whi[i], (size_t)s->KPN * s->ldw) ALLOC(s->Wlo[i], (size_t)s->KPN * s->ldw) }
        for (int i = 0; i < 2; ++i) { ALLOC(s->Hhi[i], (size_t)s->KPN * s->ldh) ALLOC(s->Hlo[i], (size_t)s->KPN * s->ldh) }
    }
    ALLOC(s->G_W, KP * KP) ALLOC(s->Minv_W, KP * KP) ALLOC(s->G_H, KP * KP) ALLOC(s->Minv_H, KP * KP)
    ALLOC(s->P_h, (size_t)s->T_h * s->maxslots_h * KP * TILE_UNITS)
    ALLOC(s->P_w, (size_t)s->T_w * s->maxslots_w * KP * TILE_UNITS)
    const size_t Wmax = (size_t)(s->nsolve_h > s->nsolve_w ? s->nsolve_h : s->nsolve_w) * SOLVE_WARPS;   // solve warps
    const size_t Gmax = (Wmax + SOLVE_GROUP - 1) / SOLVE_GROUP;
    ALLOC(s->spart, Wmax * 4) ALLOC(s->spart2, Gmax * 4)
    ALLOC(s->counters, Gmax)
    ALLOC(s->gram_part, (size_t)G * KP * KP) ALLOC(s->gram_part2, (size_t)(G / GRAM_GROUP + 1) * KP * KP)
    ALLOC(s->gram_counter, (size_t)(G / GRAM_GROUP + 2))
    ALLOC(s->ctrl, 1)
    s->max_trace = 4096;
    ALLOC(s->trace_dev, (size_t)s->max_trace * 3)
    if (shard) {
        s->xchg_cnt = (size_t)s->T_h * KP * TILE_UNITS + KP * KP;
        ALLOC(s->xchg, s->xchg_cnt)
        ALLOC(s->xscal, 4)
        ALLOC(s->peers_dev, 1)
        {
            // the shared block is a cudaMalloc of its own (not pooled): it is exported through cudaIpc
            // one exchange buffer per rank and parity for up to 8 ranks (a node), so that the fold kernels can push;
            // MFB_SHARD_SLOTS=1 keeps the pull layout (A/B runs; also used when the slots would not fit)
            s->shard_slots = getenv("MFB_SHARD_SLOTS") ? atoi(getenv("MFB_SHARD_SLOTS")) : 8;
            if (s->shard_slots < 1 || s->shard_slots > SHARD_MAXP) s->shard_slots = 1;
            if (2.0 * s->shard_slots * (double)s->xchg_cnt * sizeof(double) > 2e9) s->shard_slots = 1;
            const size_t bytes = SHARD_OFF_XCHG + (2 * (size_t)s->shard_slots * s->xchg_cnt + 2 * SHARD_GRAM_DOUBLES) * sizeof(double);
            if (cudaMalloc((void **)&s->shared, bytes) != cudaSuccess) { cudaGetLastError(); s->shared = nullptr; mfb_als_end(s); mfb_set_error("cudaMalloc of the shared exchange block failed"); return MFB_ERR_ALLOC; }
            MFB_CUDA(cudaMemsetAsync(s->shared, 0, bytes, ctx->stream));
        }
        // first collective: ||A||^2 of the whole matrix, and one verdict on negative / NA entries for all ranks
        double *hs = (double *)ctx->pinned;
        hs[0] = stats[2]; hs[1] = stats[0] < 0 ? 1.0 : 0.0; hs[2] = stats[3] > 0 ? 1.0 : 0.0; hs[3] = 0.0;
        MFB_CUDA(cudaMemcpyAsync(s->xscal, hs, 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (allreduce(user, s->xscal, 4, (void *)ctx->stream) != 0) {
            mfb_als_end(s);
            mfb_set_error("mfb_als_begin_sharded: the all-reduce callback failed");
            return MFB_ERR_CUDA;
        }
        MFB_CUDA(cudaMemcpyAsync(hs, s->xscal, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        stats[2] = hs[0];
        if (hs[2] > 0) { mfb_als_end(s); mfb_set_error("als_nmf: matrix has NA/NaN entries"); return MFB_ERR_ARG; }
        if (hs[1] > 0) { mfb_als_end(s); mfb_set_error("Negative values are not allowed"); return MFB_ERR_NEGATIVE; }
    }
#undef ALLOC
    // tensor maps over A: dims {m rows (contiguous), n cols}
    if ((rc = encode_map(&s->tm_h, A->d, A->m, A->n, A->lda, s->f32 ? 32 : 16, 128, s->f32)) != MFB_OK ||
        (rc = encode_map_rowgroups(&s->tm_w, A->d, A->n, A->lda, s->f32)) != MFB_OK) {
        mfb_als_end(s);
        return rc;
    }
    // tensor maps over the factor buffers [KP][ld]: dims {ld (contiguous), KP}, box {16, KP}
    {
        double *wb[3] = {s->Wb[0], s->Wb[1], s->Wr};
        for (int i = 0; i < 3; ++i)
            if ((rc = encode_map(&s->tm_W[i], wb[i], s->ldw, s->KP, s->ldw, 16, s->KP)) != MFB_OK) { mfb_als_end(s); return rc; }
        for (int i = 0; i < 2; ++i)
            if ((rc = encode_map(&s->tm_H[i], s->Hb[i], s->ldh, s->KP, s->ldh, 16, s->KP)) != MFB_OK) { mfb_als_end(s); return rc; }
    }
    if (s->tf32) {
        // hi / lo FP32 copies of the factors as B operands: dims {ld (contiguous), KPN}, box {32, KPN} = [KPN][128 bytes]
        for (int i = 0; i < 3; ++i)
            if ((rc = encode_map(&s->tmf_Whi[i], s->Whi[i], s->ldw, s->KPN, s->ldw, 32, s->KPN, true)) != MFB_OK ||
                (rc = encode_map(&s->tmf_Wlo[i], s->Wlo[i], s->ldw, s->KPN, s->ldw, 32, s->KPN, true)) != MFB_OK) { mfb_als_end(s); return rc; }
        for (int i = 0; i < 2; ++i)
            if ((rc = encode_map(&s->tmf_Hhi[i], s->Hhi[i], s->ldh, s->KPN, s->ldh, 32, s->KPN, true)) != MFB_OK ||
                (rc = encode_map(&s->tmf_Hlo[i], s->Hlo[i], s->ldh, s->KPN, s->ldh, 32, s->KPN, true)) != MFB_OK) { mfb_als_end(s); return rc; }
    }
    // initial factors
    cudaStream_t st = ctx->stream;
    MFB_CUDA(cudaMemcpy2DAsync(s->Wb[0], s->ldw * sizeof(double), W0, s->m * sizeof(double), s->m * sizeof(double), k,
                               cudaMemcpyHostToDevice, st));
    if (s->tf32) {
        als_split_factor_kernel<<<2 * ctx->sm_count, 256, 0, st>>>(s->Wb[0], s->ldw, s->KP, s->Whi[0], s->Wlo[0], nullptr);
        MFB_CUDA(cudaGetLastError());
    }
    if ((rc = als_upload_H(s, Hold0, s->Hb[0])) != MFB_OK) { mfb_als_end(s); return rc; }
    AlsCtrl c;
    memset(&c, 0, sizeof(c));
    c.resid_old = resid_old0;
    c.normA2 = stats[2];
    c.mn = (double)s->m_global * (double)s->n;
    c.mk = (double)s->m_global * k;
    c.kn = (double)k * s->n;
    c.max_trace = s->max_trace;
    MFB_CUDA(cudaMemcpyAsync(s->ctrl, &c, sizeof(c), cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    *out = s;
    return MFB_OK;
}

int mfb_als_begin(mfb_context *ctx, mfb_matrix *A, int k, const double *W0, const double *Hold0, double resid_old0,
                  mfb_als **out) {
    return als_begin_impl(ctx, A, A ? A->m : 0, k, W0, Hold0, resid_old0, nullptr, nullptr, out);
}

int mfb_als_begin_sharded(mfb_context *ctx, mfb_matrix *A_rows, int64_t m_global, int k, const double *W0_rows,
                          const double *Hold0, double resid_old0, mfb_allreduce_fn allreduce, void *user, mfb_als **out) {
    MFB_ARG(allreduce != nullptr, "NULL all-reduce callback");
    return als_begin_impl(ctx, A_rows, m_global, k, W0_rows, Hold0, resid_old0, allreduce, user, out);
}

int mfb_als_shard_export(mfb_als *s, void *handle64) {
    MFB_ARG(s && handle64 && s->shard && s->shared, "not a row-sharded session");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    MFB_CUDA(cudaStreamSynchronize(s->ctx->stream));          // the block is zeroed before anybody can signal into it
    cudaIpcMemHandle_t h;
    MFB_CUDA(cudaIpcGetMemHandle(&h, s->shared));
    memcpy(handle64, &h, 64);
    return MFB_OK;
}

int mfb_als_shard_connect(mfb_als *s, int rank, int world, const void *handles) {
    MFB_ARG(s && s->shard && s->shared && handles, "not a row-sharded session");
    MFB_ARG(world >= 1 && world <= SHARD_MAXP && rank >= 0 && rank < world, "rank / world out of range (at most 16 ranks)");
    MFB_ARG(s->peer_maps.empty() && s->epoch == 0, "connect once, before the first iteration");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    ShardPeers &sp = s->peers_host;
    memset(&sp, 0, sizeof(sp));
    sp.world = world; sp.rank = rank; sp.cnt = (long long)s->xchg_cnt;
    sp.slots = s->shard_slots;
    sp.push = (s->shard_slots > 1 && world <= s->shard_slots) ? 1 : 0;
    for (int q = 0; q < world; ++q) {
        if (q == rank) { sp.base[q] = s->shared; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)q * 64, 64);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (void *m : s->peer_maps) cudaIpcCloseMemHandle(m);
            s->peer_maps.clear();
            sp.world = 0;
            mfb_set_error("cudaIpcOpenMemHandle for rank %d failed: %s (the all-reduce callback stays in use)", q,
                          cudaGetErrorString(e));
            return MFB_ERR_UNSUPPORTED;
        }
        s->peer_maps.push_back(ptr);
        sp.base[q] = (unsigned char *)ptr;
    }
    MFB_CUDA(cudaMemcpyAsync(s->peers_dev, &sp, sizeof(sp), cudaMemcpyHostToDevice, s->ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return MFB_OK;
}

// development aid: globaltimer stamps (ns) of the last fold launch: [0] last block elected, [1] peers arrived,
// [4] earliest block past its stores (reset by the caller)
int mfb_als_shard_stamps(mfb_als *s, unsigned long long out[8], int reset) {
    MFB_ARG(s && s->shard && s->shared && out, "not a row-sharded session");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    MFB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    MFB_CUDA(cudaMemcpy(out, s->shared + SHARD_OFF_STAMPS, 64, cudaMemcpyDeviceToHost));
    if (reset) MFB_CUDA(cudaMemset(s->shared + SHARD_OFF_STAMPS, 0xff, 64));
    return MFB_OK;
}

// development aid: the kernel-start stamps of the last iteration (see g_tl); enable != 0 switches recording on
int mfb_als_timeline(unsigned long long out[32], int enable) {
    if (out) MFB_CUDA(cudaMemcpyFromSymbol(out, g_tl, 32 * sizeof(unsigned long long)));
    MFB_CUDA(cudaMemcpyToSymbol(g_tl_on, &enable, sizeof(int)));
    return MFB_OK;
}

int mfb_als_shard_disconnect(mfb_als *s) {
    MFB_ARG(s && s->shard, "not a row-sharded session");
    MFB_ARG(s->epoch == 0, "disconnect before the first iteration");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    for (void *m : s->peer_maps) cudaIpcCloseMemHandle(m);
    s->peer_maps.clear();
    s->peers_host.world = 0;
    return MFB_OK;
}

int mfb_als_restart(mfb_als *s, const double *Wnew) {
    MFB_ARG(s && Wnew, "NULL argument");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    MFB_CUDA(cudaMemcpy2DAsync(s->Wr, s->ldw * sizeof(double), Wnew, s->m * sizeof(double), s->m * sizeof(double), s->k,
                               cudaMemcpyHostToDevice, st));
    if (s->tf32) {
        als_split_factor_kernel<<<2 * s->ctx->sm_count, 256, 0, st>>>(s->Wr, s->ldw, s->KP, s->Whi[2], s->Wlo[2], nullptr);
        MFB_CUDA(cudaGetLastError());
    }
    MFB_CUDA(cudaStreamSynchronize(st));
    s->restart_pending = true;
    return MFB_OK;
}

// h_which: 0 = the whole iteration; 1 = only the stream kernel of the H-step (snmf: run ahead for the convergence
// test); 2 = the iteration without that stream kernel (it has run ahead)
// parameter blocks of the two half-steps of the iteration that reads W / H from buffer `cur` and writes `cur ^ 1`
static void half_params(const mfb_als *s, int cur, bool use_restart_w, HalfParams &h, HalfParams &w) {
    memset(&h, 0, sizeof(h));
    const int nxt = cur ^ 1;
    if (s->snmf) h.reg_all = s->beta;     // H-step: (W'W + beta 11') H = W'A
    // H-step: fixed W (or the restart draw), solve H
    h.Fin = use_restart_w ? s->Wr : s->Wb[cur]; h.ldin = s->ldw; h.nfix = s->m;
    h.Fout = s->Hb[nxt]; h.Fold = s->Hb[cur]; h.ldout = s->ldh;
    h.G = s->G_W; h.Minv = s->Minv_W; h.gram_part = s->gram_part; h.gram_part2 = s->gram_part2; h.gram_counter = s->gram_counter;
    h.P = s->P_h; h.maxslots = s->maxslots_h; h.group_counter = s->counters;
    h.spart = s->spart; h.spart2 = s->spart2;
    h.ctrl = s->ctrl; h.trace = s->trace_dev;
    h.k = s->k; h.T = s->T_h; h.SPT = s->SPT_h; h.U = (long long)s->T_h * s->SPT_h; h.nvalid = s->n;
    h.Gstream = s->ctx->sm_count;
    h.shard = s->shard ? 1 : 0; h.xscal = s->xscal;
    h.invert_later = (!s->shard && s->KB >= 7 && !getenv("MFB_INVERT_IN_STREAM")) ? 1 : 0;
    h.peers = als_peer_mode(s) ? s->peers_dev : nullptr;
    h.peer_invert_later = (h.peers && !getenv("MFB_SHARD_INVERT_IN_STREAM")) ? 1 : 0;
    // W-step: fixed H (just solved), solve W
    w = h;
    w.peer_invert_later = 0;
    w.Fin = s->Hb[nxt]; w.ldin = s->ldh; w.nfix = s->n;
    w.Fout = s->Wb[nxt]; w.Fold = s->Wb[cur]; w.ldout = s->ldw;
    w.G = s->G_H; w.Minv = s->Minv_H;
    w.P = s->P_w; w.maxslots = s->maxslots_w;
    w.T = s->T_w; w.SPT = s->SPT_w; w.U = (long long)s->T_w * s->SPT_w; w.nvalid = s->m;
    if (s->snmf) { w.reg_all = 0.0; w.reg_diag = s->eta * s->eta; }     // W-step: (HH' + eta^2 I) W' = H A'
}

// h_which: 0 = the whole iteration; 1 = only the stream kernel of the H-step (snmf: run ahead for the convergence
// test); 2 = the iteration without that stream kernel (it has run ahead)
static int enqueue_iteration(mfb_als *s, int cur, bool use_restart_w, int h_which = 0) {
    HalfParams h, w;
    half_params(s, cur, use_restart_w, h, w);
    const int nxt = cur ^ 1;
    const bool peer = als_peer_mode(s);
    const bool pdl = als_use_pdl(s);
    if (s->shard) h.epoch = w.epoch = ++s->epoch;
    const CUtensorMap &tmW = use_restart_w ? s->tm_W[2] : s->tm_W[cur];
    Tf32Extra xh, xw;
    if (s->tf32) {
        const int wi = use_restart_w ? 2 : cur;
        xh.bhi = &s->tmf_Whi[wi]; xh.blo = &s->tmf_Wlo[wi]; xh.out_hi = s->Hhi[nxt]; xh.out_lo = s->Hlo[nxt];
        xw.bhi = &s->tmf_Hhi[nxt]; xw.blo = &s->tmf_Hlo[nxt]; xw.out_hi = s->Whi[nxt]; xw.out_lo = s->Wlo[nxt];
    }
    if (!s->shard) {
        MFB_TRY(launch_half_dispatch(s, false, h, tmW, h_which, xh));
        if (h_which == 1) return MFB_OK;
    } else {
        // local W_r'A_r and W_r'W_r -> one buffer -> sum over the ranks -> invert -> NNLS (replicated, identical H)
        cudaStream_t st = s->ctx->stream;
        const size_t cnt = (size_t)s->T_h * s->KP * TILE_UNITS + (size_t)s->KP * s->KP;
        MFB_TRY(launch_half_dispatch(s, false, h, tmW, 1, xh));
        // peer memory: (W'W)^-1 is formed from the published W_q'W_q after the stream kernel -- by an extra block of
        // the fold kernel (ranks up to 32) or by the kernel that closes the previous iteration (1024 threads, above)
        const bool inv_in_fold = peer && h.peer_invert_later && s->KP <= 32;
        if (peer && !inv_in_fold) {
            if (h.peer_invert_later) MFB_CUDA(launch_pdl(als_shard_close_invert_kernel, 1, 1024, 0, st, pdl, h, s->KP));
            else MFB_CUDA(launch_pdl(als_shard_close_kernel, 1, 32, 0, st, pdl, h));
        }
        double *xbuf = peer ? (double *)(s->shared + SHARD_OFF_XCHG) + (size_t)(h.epoch & 1) * s->shard_slots * s->xchg_cnt : s->xchg;
        // (ranks up to 32: the fold launch carries two more blocks, one for (W'W)^-1 and one that closes the previous iteration)
        MFB_CUDA(launch_pdl(als_shard_fold_kernel, s->T_h * s->KP + (inv_in_fold ? 2 : 0), TILE_UNITS, 0, st, pdl, h, s->KP, xbuf,
                            inv_in_fold ? s->T_h * s->KP : -1, inv_in_fold ? s->T_h * s->KP + 1 : -1));
        MFB_TRY(prof_mark(s));
        if (!peer) {
            if (s->allreduce(s->allreduce_user, s->xchg, (int64_t)cnt, (void *)st) != 0) {
                mfb_set_error("mfb_als_iterate: the all-reduce callback failed");
                return MFB_ERR_CUDA;
            }
            MFB_CUDA(launch_pdl(als_shard_invert_kernel, 1, 1024, 0, st, pdl, h, s->KP,
                                (const double *)(s->xchg + (size_t)s->T_h * s->KP * TILE_UNITS), 3));
            MFB_TRY(prof_mark(s));
        }
        HalfParams hs = h;
        hs.P = s->xchg; hs.maxslots = 1;
        MFB_TRY(launch_half_dispatch(s, false, hs, tmW, 2, xh));
    }
    MFB_TRY(launch_half_dispatch(s, true, w, s->tm_H[nxt], 0, xw));
    if (s->shard && !peer) {
        cudaStream_t st = s->ctx->stream;
        if (s->allreduce(s->allreduce_user, s->xscal, 4, (void *)st) != 0) {
            mfb_set_error("mfb_als_iterate: the all-reduce callback failed");
            return MFB_ERR_CUDA;
        }
        MFB_CUDA(launch_pdl(als_shard_finish_kernel, 1, 32, 0, st, pdl, w));
        MFB_TRY(prof_mark(s));
    }
    return MFB_OK;
}

int mfb_als_iterate(mfb_als *s, int max_new_iters, double eps_conv, int fixed_iters, int *iters_total, int *converged) {
    return mfb_als_iterate_draws(s, max_new_iters, eps_conv, fixed_iters, iters_total, converged, nullptr);
}

int mfb_als_iterate_draws(mfb_als *s, int max_new_iters, double eps_conv, int fixed_iters, int *iters_total, int *converged,
                          long long *maxcol_draws) {
    MFB_ARG(s, "NULL session");
    const bool count = maxcol_draws != nullptr && s->KP <= 32 && !s->shard;
    if (maxcol_draws) *maxcol_draws = count ? 0 : -1;
    unsigned long long draws0 = 0, draws1 = 0;
    bool first_batch = true;
    MFB_ARG(eps_conv > 0, "ALS-NMF: Invalid argument 'eps_conv' - value should be positive");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    if (converged) *converged = 0;
    int status = MFB_OK;
    int remaining = max_new_iters;
    AlsCtrl *hc = (AlsCtrl *)s->ctx->pinned;
    while (remaining > 0) {
        const int batch = fixed_iters ? remaining : (remaining < 8 ? remaining : 8);
        // arm the control block for this batch (later batches of the call re-arm the copy read back at the end of the
        // previous one: nothing else writes the block in between)
        if (first_batch) {
            MFB_CUDA(cudaMemcpyAsync(hc, s->ctrl, sizeof(AlsCtrl), cudaMemcpyDeviceToHost, st));
            MFB_CUDA(cudaStreamSynchronize(st));
        }
        hc->stop = 0; hc->reason = 0; hc->iters_done = 0; hc->fixed = fixed_iters ? 1 : 0; hc->eps = eps_conv;
        hc->groups_done[0] = hc->groups_done[1] = 0; hc->nzmask = 0ull;
        hc->count_draws = count ? 1 : 0;
        hc->pending_epoch = 0; hc->unpublished_epoch = 0;
        if (first_batch) { draws0 = hc->draws; first_batch = false; }
        if (batch > s->max_trace) { mfb_set_error("too many iterations in one batch"); return MFB_ERR_ARG; }
        MFB_CUDA(cudaMemcpyAsync(s->ctrl, hc, sizeof(AlsCtrl), cudaMemcpyHostToDevice, st));
        int cur = s->cur;
        for (int b = 0; b < batch; ++b) {
            MFB_TRY(enqueue_iteration(s, cur, s->restart_pending && b == 0));
            cur ^= 1;
        }
        if (als_peer_mode(s)) {          // the last iteration of the batch is still open
            HalfParams hc2;
            memset(&hc2, 0, sizeof(hc2));
            hc2.ctrl = s->ctrl; hc2.trace = s->trace_dev; hc2.peers = s->peers_dev;
            MFB_CUDA(launch_pdl(als_shard_close_kernel, 1, 32, 0, st, als_use_pdl(s), hc2));
        }
        MFB_CUDA(cudaMemcpyAsync(hc, s->ctrl, sizeof(AlsCtrl), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        MFB_TRY(check_mbar_abort("mfb_als_iterate"));
        draws1 = hc->draws;
        const int done = hc->iters_done;
        if (done > 0) {
            const size_t old = s->trace.size();
            s->trace.resize(old + (size_t)done * 3);
            MFB_CUDA(cudaMemcpy(s->trace.data() + old, s->trace_dev, (size_t)done * 3 * sizeof(double), cudaMemcpyDeviceToHost));
            s->resid = hc->resid;
            s->restart_pending = false;
        }
        s->iters_total += done;
        if (done & 1) s->cur ^= 1;
        remaining -= done;
        if (hc->stop) {
            if (hc->reason == 3 && remaining <= 0) {
                // W'W of the final W is singular, but R never runs another H-step: not a failure
            } else if (hc->reason == 1) {
                if (converged) *converged = 1;
            } else {
                // the iteration that was running failed: R has already done `i <- i + 1`
                s->iters_total += 1;
                const double qnan = NAN;
                s->trace.push_back(qnan); s->trace.push_back(qnan); s->trace.push_back(qnan);
                status = (hc->reason == 2) ? MFB_ZERO_ROW : MFB_SINGULAR;
            }
            break;
        }
        if (done == 0) break;   // defensive: nothing ran
    }
    if (iters_total) *iters_total = s->iters_total;
    if (count) *maxcol_draws = (long long)(draws1 - draws0);
    return status;
}

// ---------------------------------------------------------------------------
// snmf sessions: NMF::nmf(method = "snmf/r") (the call at R/bicluster.R:363) on the ALS kernels
// ---------------------------------------------------------------------------
// W = apply(W, 2, function(x) x / sqrt(sum(x^2)))
static void unit_columns(const double *W, int64_t m, int k, std::vector<double> &out) {
    out.resize((size_t)m * k);
    for (int c = 0; c < k; ++c) {
        long double ss = 0.0L;                      // R's sum() accumulates in long double
        for (int64_t i = 0; i < m; ++i) ss += (long double)(W[(size_t)c * m + i] * W[(size_t)c * m + i]);
        const double nrm = sqrt((double)ss);
        for (int64_t i = 0; i < m; ++i) out[(size_t)c * m + i] = W[(size_t)c * m + i] / nrm;
    }
}

static int snmf_reset_test_state(mfb_als *s) {
    cudaStream_t st = s->ctx->stream;
    for (int i = 0; i < 2; ++i) {
        MFB_CUDA(cudaMemsetAsync(s->idxW[i], 0, (size_t)s->m * sizeof(int), st));   // idxWold = rep(0, m)
        MFB_CUDA(cudaMemsetAsync(s->idxH[i], 0, (size_t)s->n * sizeof(int), st));   // idxHold = rep(0, n)
    }
    s->idx_cur = 0;
    s->inc = 0;
    s->erravg1 = -1.0;
    s->prestreamed = false;
    return MFB_OK;
}

int mfb_snmf_begin(mfb_context *ctx, mfb_matrix *A, int k, const double *W0, double beta, double eta, mfb_als **out) {
    MFB_ARG(ctx && A && W0 && out, "NULL argument");
    MFB_ARG(A->elem_bytes == 8, "snmf needs the FP64 resident matrix");
    MFB_ARG(beta >= 0.0, "snmf: beta must be >= 0");
    double stats[5];
    MFB_TRY(mfb_matrix_stats(A, stats));
    if (eta < 0.0) eta = stats[1];                  // .nmf_snmf: if (eta < 0) eta = max(A)
    MFB_ARG(k >= 1 && k <= 64, "k must be in 1..64");
    std::vector<double> Wn, Hz((size_t)k * A->n, 0.0);
    unit_columns(W0, A->m, k, Wn);
    mfb_als *s = nullptr;
    MFB_TRY(als_begin_impl(ctx, A, A->m, k, Wn.data(), Hz.data(), stats[1], nullptr, nullptr, &s));
    s->snmf = true;
    s->beta = beta;
    s->eta = eta;
    int rc;
#define ALLOC(ptr, cnt) if ((rc = dev_alloc(s, &(ptr), (cnt))) != MFB_OK) { mfb_als_end(s); return rc; }
    ALLOC(s->idxW[0], (size_t)s->m) ALLOC(s->idxW[1], (size_t)s->m) ALLOC(s->idxH[0], (size_t)s->n) ALLOC(s->idxH[1], (size_t)s->n)
    ALLOC(s->chk_part, (size_t)(s->T_h + s->T_w) * 8)
#undef ALLOC
    if ((rc = snmf_reset_test_state(s)) != MFB_OK) { mfb_als_end(s); return rc; }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = s;
    return MFB_OK;
}

int mfb_snmf_restart(mfb_als *s, const double *Wnew) {
    MFB_ARG(s && Wnew && s->snmf, "not an snmf session");
    std::vector<double> Wn;
    unit_columns(Wnew, s->m, s->k, Wn);
    MFB_TRY(mfb_als_restart(s, Wn.data()));
    return snmf_reset_test_state(s);
}

// the two launches of the convergence test on the factors in buffer `cur`; with_grad = 0: only the sums of squares
static int snmf_enqueue_check(mfb_als *s, int cur, int with_grad) {
    // half-step blocks of the iteration that PRODUCED buffer `cur` (its W-step: G_H, P_w are still in place) ...
    HalfParams hp, wp, hn, wn;
    half_params(s, cur ^ 1, false, hp, wp);
    // ... and of the iteration that will READ it (its H-step's stream kernel has run ahead: G_W, P_h)
    half_params(s, cur, false, hn, wn);
    hn.Fout = s->Hb[cur];              // the factor under test, not the one the next H-step will write
    cudaStream_t st = s->ctx->stream;
    const int oi = s->idx_cur, ni = oi ^ 1;
    const size_t smem = ((size_t)s->KP * s->KP + (size_t)s->KP * TILE_UNITS) * sizeof(double);
    MFB_CUDA(cudaFuncSetAttribute(snmf_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 blk(TILE_UNITS, SNMF_CHECK_Y);
    snmf_check_kernel<<<s->T_h, blk, smem, st>>>(hn, s->KP, with_grad, s->idxH[oi], s->idxH[ni], s->chk_part);
    snmf_check_kernel<<<s->T_w, blk, smem, st>>>(wp, s->KP, with_grad, s->idxW[oi], s->idxW[ni], s->chk_part + (size_t)s->T_h * 8);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int mfb_snmf_iterate(mfb_als *s, int max_new_iters, int wminchange, int iconv, double eps_conv, int *iters_total,
                     int *converged, double *objective) {
    MFB_ARG(s && s->snmf, "not an snmf session");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    if (converged) *converged = 0;
    int status = MFB_OK;
    int remaining = max_new_iters;
    bool stopped = false, sums_current = false, have_ctrl = false;
    AlsCtrl *hc = (AlsCtrl *)s->ctx->pinned;
    const size_t npart = (size_t)(s->T_h + s->T_w) * 8;
    std::vector<double> part_vec;
    double *part_h = (double *)((char *)s->ctx->pinned + 4096);                       // pinned, after the control block
    if (4096 + npart * sizeof(double) > s->ctx->pinned_bytes) { part_vec.resize(npart); part_h = part_vec.data(); }
    double tot[2][5];
    auto read_check = [&](bool copied) -> int {
        if (!copied) {
            MFB_CUDA(cudaMemcpyAsync(part_h, s->chk_part, npart * sizeof(double), cudaMemcpyDeviceToHost, st));
            MFB_CUDA(cudaStreamSynchronize(st));
        }
        for (int f = 0; f < 2; ++f)
            for (int q = 0; q < 5; ++q) {
                double a = 0.0;
                const int b0 = f ? s->T_h : 0, b1 = f ? s->T_h + s->T_w : s->T_h;
                for (int b = b0; b < b1; ++b) a += part_h[(size_t)b * 8 + q];     // block order
                tot[f][q] = a;
            }
        return MFB_OK;
    };
    auto set_objective = [&]() {
        // .snmf.objective = 1/2 (||A - WH||^2 + eta sum(W^2) + beta sum(colSums(H)^2)); the W-step solve kernel's scalars
        // give  ||A||^2 - 2 sum w.(AH') + sum w'(HH' + eta^2 I) w  = resid^2 m n
        const double mn = (double)s->m * (double)s->n;
        const double fit = s->resid * s->resid * mn - s->eta * s->eta * tot[1][3];
        s->objective = 0.5 * (fit + s->eta * tot[1][3] + s->beta * tot[0][4]);
    };
    while (remaining > 0 && !stopped) {
        const bool first = s->erravg1 < 0.0;                       // length(erravg1) == 0
        int run = first ? 1 : 5 - (s->iters_total % 5);
        if (run > remaining) run = remaining;
        const bool check = first || ((s->iters_total + run) % 5 == 0);
        if (!have_ctrl) {      // (later batches of this call re-arm the copy they read back at the end of the previous one)
            MFB_CUDA(cudaMemcpyAsync(hc, s->ctrl, sizeof(AlsCtrl), cudaMemcpyDeviceToHost, st));
            MFB_CUDA(cudaStreamSynchronize(st));
            have_ctrl = true;
        }
        hc->stop = 0; hc->reason = 0; hc->iters_done = 0; hc->fixed = 1; hc->eps = 1.0;
        hc->groups_done[0] = hc->groups_done[1] = 0; hc->nzmask = 0ull;
        hc->count_draws = 0; hc->pending_epoch = 0; hc->unpublished_epoch = 0;
        MFB_CUDA(cudaMemcpyAsync(s->ctrl, hc, sizeof(AlsCtrl), cudaMemcpyHostToDevice, st));
        int cur = s->cur;
        for (int b = 0; b < run; ++b) {
            MFB_TRY(enqueue_iteration(s, cur, s->restart_pending && b == 0, (b == 0 && s->prestreamed) ? 2 : 0));
            cur ^= 1;
        }
        s->prestreamed = false;
        sums_current = false;
        if (check) {
            MFB_TRY(enqueue_iteration(s, cur, false, 1));           // W'A and W'W + beta 11' of the new W
            MFB_TRY(snmf_enqueue_check(s, cur, 1));
            MFB_CUDA(cudaMemcpyAsync(part_h, s->chk_part, npart * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
        MFB_CUDA(cudaMemcpyAsync(hc, s->ctrl, sizeof(AlsCtrl), cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        MFB_TRY(check_mbar_abort("mfb_snmf_iterate"));
        const int done = hc->iters_done;
        if (done > 0) {
            s->resid = hc->resid;
            s->restart_pending = false;
        }
        s->iters_total += done;
        if (done & 1) s->cur ^= 1;
        remaining -= done;
        if (done < run) {
            // the iteration that was running failed (`i <- i + 1` precedes the failure in R as well)
            s->iters_total += 1;
            status = (hc->reason == 2) ? MFB_ZERO_ROW : MFB_SINGULAR;
            break;
        }
        if (check) {
            MFB_TRY(read_check(true));
            sums_current = true;
            const double conv = tot[0][0] + tot[1][0], convnum = tot[0][1] + tot[1][1];
            const double erravg = conv / convnum;
            const long long changedW = (long long)tot[1][2], changedH = (long long)tot[0][2];
            s->inc = (changedW <= wminchange && changedH == 0) ? s->inc + 1 : 0;
            if (s->erravg1 < 0.0) s->erravg1 = (erravg == erravg) ? erravg : 0.0;
            if (getenv("MFB_SNMF_DEBUG"))
                fprintf(stderr, "snmf test i=%d inc=%d changedW=%lld changedH=%lld convH=%.17g convW=%.17g numH=%.0f numW=%.0f erravg=%.6e erravg1=%.6e\n",
                        s->iters_total, s->inc, changedW, changedH, tot[0][0], tot[1][0], tot[0][1], tot[1][1], erravg, s->erravg1);
            if (s->inc >= iconv && erravg <= eps_conv * s->erravg1) {
                stopped = true;
                if (converged) *converged = 1;
            } else {
                s->idx_cur ^= 1;                                    // idxWold = idxW; idxHold = idxH
                if (hc->stop && hc->reason == 3) {
                    // W'W + beta 11' of the new W is singular: the next iteration's .fcnnls fails in solve()
                    if (remaining > 0) { s->iters_total += 1; status = MFB_SINGULAR; break; }
                } else {
                    s->prestreamed = true;
                }
            }
        }
    }
    if (status == MFB_OK && objective) {
        if (!sums_current) {
            MFB_TRY(snmf_enqueue_check(s, s->cur, 0));
            MFB_TRY(read_check(false));
        }
        set_objective();
        *objective = s->objective;
    }
    if (iters_total) *iters_total = s->iters_total;
    return status;
}

static void tf32_timing_begin() {
    if (!getenv("MFB_TF32_TIMING")) return;
    unsigned long long z[16] = {0};
    int one = 1;
    cudaMemcpyToSymbol(g_tf32_ticks, z, sizeof(z));
    cudaMemcpyToSymbol(g_tf32_timing, &one, sizeof(int));
}
static void tf32_timing_end(int iters) {
    if (!getenv("MFB_TF32_TIMING")) return;
    unsigned long long t[16];
    int zero = 0;
    cudaMemcpyFromSymbol(t, g_tf32_ticks, sizeof(t));
    cudaMemcpyToSymbol(g_tf32_timing, &zero, sizeof(int));
    fprintf(stderr, "tf32 converter (block 0, thread 0) clocks per iteration: loop %llu | wait full %llu | lds+split %llu | wait ta_empty %llu | sttm %llu | - %llu | wait d_full %llu | ldtm+fp64 %llu\n",
            t[0] / iters, t[1] / iters, t[2] / iters, t[3] / iters, t[4] / iters, t[5] / iters, t[6] / iters, t[7] / iters);
}

int mfb_als_profile(mfb_als *s, int iters, double *ms) {
    MFB_ARG(s && ms && iters >= 1, "bad argument");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    std::vector<cudaEvent_t> ev;
    cudaEvent_t e0;
    MFB_CUDA(cudaEventCreate(&e0));
    MFB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    MFB_CUDA(cudaEventRecord(e0, s->ctx->stream));
    ev.push_back(e0);
    s->prof = &ev;
    int it = 0, conv = 0;
    tf32_timing_begin();
    const int rc = mfb_als_iterate(s, iters, 1e-4, 1, &it, &conv);
    s->prof = nullptr;
    MFB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    tf32_timing_end(iters);
    if (getenv("MFB_GRAM_TIMING")) {
        unsigned long long st[16];
        MFB_CUDA(cudaMemcpyFromSymbol(st, g_gram_stamps, sizeof(st)));
        fprintf(stderr, "gram side job of the inverting CTA (last launch, us): partial %.1f | wait+l1 start %.1f | l1 fold -> l2 start %.1f | l2 fold %.1f | inverse %.1f\n",
                (st[1] - st[0]) * 1e-3, (st[2] - st[1]) * 1e-3, (st[3] - st[2]) * 1e-3, (st[4] - st[3]) * 1e-3, (st[5] - st[4]) * 1e-3);
    }
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t nk = ev.size() - 1;
    // sharded, callback exchange: H stream, fold, invert, H solve, W stream, W solve, finish (collectives in between);
    // sharded, peer memory: H stream, fold (+ wait for the peers, invert), H solve, W stream, W solve (+ wait, finish)
    const size_t per = s->shard ? (als_peer_mode(s) ? 5 : 7) : 4;
    for (size_t i = 0; i < nk; ++i) {
        float t = 0.f;
        MFB_CUDA(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
        acc[i % per] += t;
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (rc != MFB_OK) return rc;
    if (nk != per * iters) { mfb_set_error("mfb_als_profile: expected %d kernel launches, saw %zu", (int)per * iters, nk); return MFB_ERR_CUDA; }
    for (size_t q = 0; q < per; ++q) ms[q] = acc[q] / iters;
    return MFB_OK;
}

int mfb_als_result(mfb_als *s, double *W, double *H, double *resid, double *trace) {
    MFB_ARG(s, "NULL session");
    MFB_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    if (W)
        MFB_CUDA(cudaMemcpy2DAsync(W, s->m * sizeof(double), s->Wb[s->cur], s->ldw * sizeof(double), s->m * sizeof(double),
                                   s->k, cudaMemcpyDeviceToHost, st));
    if (H) {
        std::vector<double> tmp((size_t)s->k * s->n);
        MFB_CUDA(cudaMemcpy2DAsync(tmp.data(), s->n * sizeof(double), s->Hb[s->cur], s->ldh * sizeof(double),
                                   s->n * sizeof(double), s->k, cudaMemcpyDeviceToHost, st));
        MFB_CUDA(cudaStreamSynchronize(st));
        for (int64_t j = 0; j < s->n; ++j)
            for (int c = 0; c < s->k; ++c) H[(size_t)j * s->k + c] = tmp[(size_t)c * s->n + j];
    }
    MFB_CUDA(cudaStreamSynchronize(st));
    if (resid) *resid = s->resid;
    if (trace && !s->trace.empty()) memcpy(trace, s->trace.data(), s->trace.size() * sizeof(double));
    return MFB_OK;
}

int mfb_als_run(mfb_context *ctx, mfb_matrix *A, int k, const double *W0, const double *Hold0, double resid_old0,
                int max_iter, double eps_conv, int fixed_iters, double *W, double *H, double *resid, int *iters_done,
                int *converged, double *trace) {
    mfb_als *s = nullptr;
    MFB_TRY(mfb_als_begin(ctx, A, k, W0, Hold0, resid_old0, &s));
    int it = 0, conv = 0;
    int status = mfb_als_iterate(s, max_iter, eps_conv, fixed_iters, &it, &conv);
    if (status >= 0) {
        int rc = mfb_als_result(s, W, H, resid, trace);
        if (rc != MFB_OK) status = rc;
    }
    if (iters_done) *iters_done = it;
    if (converged) *converged = conv;
    mfb_als_end(s);
    return status;
}

int mfb_als_run_host(mfb_context *ctx, const double *A_host, int64_t m, int64_t n, int k, const double *W0,
                     const double *Hold0, int max_iter, double eps_conv, int fixed_iters, double *W, double *H,
                     double *resid, int *iters_done, int *converged, double *trace) {
    mfb_matrix *A = nullptr;
    MFB_TRY(mfb_matrix_upload(ctx, A_host, m, n, &A));
    double stats[5];
    int rc = mfb_matrix_stats(A, stats);
    if (rc == MFB_OK)
        rc = mfb_als_run(ctx, A, k, W0, Hold0, stats[1], max_iter, eps_conv, fixed_iters, W, H, resid, iters_done,
                         converged, trace);
    mfb_matrix_free(A);
    return rc;
}

int mfb_nnls(mfb_context *ctx, const double *C, const double *B, int64_t r, int k, int64_t pcols, double *K) {
    return mfb_nnls_draws(ctx, C, B, r, k, pcols, K, nullptr);
}

int mfb_nnls_draws(mfb_context *ctx, const double *C, const double *B, int64_t r, int k, int64_t pcols, double *K,
                   long long *maxcol_draws) {
    MFB_ARG(ctx && C && B && K, "NULL argument");
    MFB_ARG(k >= 1 && k <= 64 && r >= 1 && pcols >= 1, "bad sizes");
    MFB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int KB = (k + 7) / 8;
    if (KB == 5) KB = 6;
    if (KB == 7) KB = 8;
    const int KP = KB * 8;
    double *dC, *dB, *dR, *dG, *dM, *dK, *part;
    int *dsing;
    unsigned long long *ddraws = nullptr;
    DevBuf buf(ctx);
    if (maxcol_draws) {
        *maxcol_draws = -1;                                  // unknown (ranks above 32 use block pivoting: no per-variable events)
        if (k <= 32) {
            int rc = buf.alloc(&ddraws, (size_t)1);
            if (rc != MFB_OK) return rc;
            MFB_CUDA(cudaMemsetAsync(ddraws, 0, sizeof(unsigned long long), st));
        }
    }
    {
        int rc = buf.alloc(&dC, (size_t)r * k);
        if (rc == MFB_OK) rc = buf.alloc(&dB, (size_t)r * pcols);
        if (rc == MFB_OK) rc = buf.alloc(&dR, (size_t)k * pcols);
        if (rc == MFB_OK) rc = buf.alloc(&dK, (size_t)k * pcols);
        if (rc == MFB_OK) rc = buf.alloc(&dG, (size_t)KP * KP);
        if (rc == MFB_OK) rc = buf.alloc(&dM, (size_t)KP * KP);
        if (rc == MFB_OK) rc = buf.alloc(&part, (size_t)64 * KP * KP);
        if (rc == MFB_OK) rc = buf.alloc(&dsing, 1);
        if (rc != MFB_OK) return rc;
    }
    MFB_CUDA(cudaMemcpyAsync(dC, C, sizeof(double) * r * k, cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaMemcpyAsync(dB, B, sizeof(double) * r * pcols, cudaMemcpyHostToDevice, st));
    int nb = (int)((r + 2047) / 2048);
    if (nb > 64) nb = 64;
    gram_partial_kernel<<<nb, 256, (size_t)k * 65 * sizeof(double), st>>>(dC, r, r, k, KP, part);
    gram_finalize_kernel<<<1, 256, (size_t)(KP * KP + 2) * sizeof(double), st>>>(part, nb, k, KP, dG, dM, dsing);
    small_crossprod_kernel<<<(unsigned)pcols, 256, 0, st>>>(dC, dB, r, k, pcols, dR);
    {
        int rc = MFB_ERR_UNSUPPORTED;
        switch (KB) {
            case 1: rc = launch_nnls_columns<1>(dG, dM, dR, k, pcols, dK, st, ddraws); break;
            case 2: rc = launch_nnls_columns<2>(dG, dM, dR, k, pcols, dK, st, ddraws); break;
            case 3: rc = launch_nnls_columns<3>(dG, dM, dR, k, pcols, dK, st, ddraws); break;
            case 4: rc = launch_nnls_columns<4>(dG, dM, dR, k, pcols, dK, st, ddraws); break;
            case 6: rc = launch_nnls_columns<6>(dG, dM, dR, k, pcols, dK, st, ddraws); break;
            case 8: rc = launch_nnls_columns<8>(dG, dM, dR, k, pcols, dK, st, ddraws); break;
        }
        if (rc != MFB_OK) return rc;
    }
    MFB_CUDA(cudaGetLastError());
    int sing = 0;
    MFB_CUDA(cudaMemcpyAsync(K, dK, sizeof(double) * k * pcols, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaMemcpyAsync(&sing, dsing, sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned long long hd = 0;
    if (ddraws) MFB_CUDA(cudaMemcpyAsync(&hd, ddraws, sizeof(hd), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    if (ddraws) *maxcol_draws = sing ? 0 : (long long)hd;    // solve(CtC) fails before .fcnnls reaches a max.col call
    return sing ? MFB_SINGULAR : MFB_OK;
}

}  // extern "C"

// FP32 path of the ALS stream kernels: single-precision tensor cores (tcgen05.mma kind::tf32), 3 x TF32 split.
//
// A resident in FP32 halves the HBM traffic of an iteration (2*m*n*4 bytes), but the contraction then needs
// 8 flop per streamed byte: beyond the FP64 tensor pipe (the FP32-storage kernel that widens to FP64 runs at
// 0.45 of the FP32 HBM roofline).  Here the products run on the 5th-generation tensor cores instead:
//
//   * every FP32 entry of a tile is split by the threads that take it out of the TMA ring into
//       a = a_hi + a_lo,  a_hi = the top 19 bits (exactly a TF32 number), a_lo = a - a_hi (exact in FP32),
//     and both parts are written to TENSOR MEMORY as the A operand (lane = unit of the tile, column = reduction
//     index) with tcgen05.st -- the tile is read from shared memory once, by the generic proxy, and the MMAs read
//     their big operand from TMEM, so shared-memory bandwidth is not the limit;
//   * the fixed factor comes pre-split (FP64 -> hi / lo FP32 arrays written by als_split_factor_kernel after every
//     solve) as the B operand: K-major [rank][32] tiles that TMA lands in the 128-byte-swizzled layout the UMMA
//     shared-memory descriptor describes;
//   * per 32-deep stage 4 x 2 MMAs (the hi and lo tiles of B are adjacent in the ring):
//       [D_main | D_corr] += A_hi [B_hi ; B_lo]'   (128 x 2 KPN x 8),     D_corr += A_lo B_hi'   (128 x KPN x 8)
//     (the a_lo b_lo products, ~2^-22 of the result, are put back as their mean, see the converter)
//     with FP32 accumulators in TMEM.  Two things shape this: (i) the tensor core adds with truncation, so an
//     accumulator is never carried across stages -- every stage gets fresh accumulators (a ring of them) which the
//     converter threads add up in registers (tcgen05.ld, one stage late, overlapping the next MMAs) as an unevaluated
//     FP32 pair (two-sum: ~48 bits, round-to-nearest), widened to FP64 only when the stream-K partial is written; the
//     NNLS and everything downstream are the FP64 path, unchanged; (ii) an MMA that accumulates into the result of
//     the previous one waits for it (~180 cycles measured at N = 16), so the products go to two independent
//     accumulator chains of depth 4 per stage instead of one of depth 12.
// Measured error of a 512-deep contraction: 1.3e-5 with one running FP32 accumulator, ~2e-7 with per-stage flushes.
//
//     (iii) ONE thread can issue a tcgen05.mma only every ~117 cycles whatever its shape (scripts/microbench/
//     umma_rate.cu: N = 16 and N = 128 cost the same; four warps issue four times as many) -- a stage's 8 small MMAs
//     would take longer than its 16 KB take to arrive from HBM, so four issuer warps (one per SM sub-partition) take
//     the stages round-robin, each with its own A-operand and accumulator buffers.
// Warp roles (768 threads): 0-3 / 4-7 two converter groups (even / odd stages: shared memory -> split -> TMEM),
// 8 TMA producer, 9-11 Gram side job (as in the FP64 kernel), 12-15 MMA issuers (stage it -> warp 12 + it % 4; warp 12
// owns the TMEM allocation), 16-19 / 20-23 two absorber groups (even / odd stages: accumulators -> registers ->
// stream-K partial, one partial per group).
#pragma once

#define TF32_THREADS 768
__device__ unsigned long long g_tf32_ticks[16];
__device__ int g_tf32_timing = 0;
__device__ int g_tf32_comp = 1;      // put the mean truncation loss of the tensor-core accumulation back (MFB_TF32_NOCOMP=1: off)
#define TF32_TICK(slot) do { if (timing) { const long long n_ = clock64(); tacc[slot] += (unsigned long long)(n_ - tmark); tmark = n_; } } while (0)
#define TF32_NAB 4                // A-operand buffers in TMEM (64 columns each: hi | lo)

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // K-major operand, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    // kind::tf32, FP32 accumulate, A and B K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tf32_group_bar() { asm volatile("bar.sync 3, 640;" ::: "memory"); }   // converters + MMA warps + absorbers
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// hi / lo FP32 copies of a factor buffer [KP][ld] (FP64): hi = the value rounded to FP32 with the low 13 mantissa
// bits rounded away (the nearest TF32 number), lo = x - hi: a SIGNED remainder, so that neither the dropped lo * lo
// products nor the tensor core's truncation of lo toward zero is biased.  Rows >= KP of the FP32 arrays stay zero.
__global__ void __launch_bounds__(256) als_split_factor_kernel(const double *__restrict__ F, int64_t ld, int KP,
                                                               float *__restrict__ hi, float *__restrict__ lo,
                                                               const AlsCtrl *__restrict__ ctrl) {
    pdl_prologue();
    if (ctrl && *(volatile const int *)&ctrl->stop) return;
    const int64_t total = (int64_t)KP * ld;
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const double x = F[e];
        const float h = __uint_as_float((__float_as_uint((float)x) + 0x1000u) & 0xffffe000u);   // nearest TF32 number
        hi[e] = h;
        lo[e] = (float)(x - (double)h);                      // signed remainder (the tensor core truncates it toward zero)
    }
}

template <int KB, bool WSTEP>
__global__ void __launch_bounds__(TF32_THREADS, 1)
als_stream_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                       const __grid_constant__ CUtensorMap tmBlo, const HalfParams p, const int nstage) {
    constexpr int KP = KB * 8;
    constexpr int KPN = (KP + 15) / 16 * 16;            // MMA N (M = 128 needs a multiple of 16); extra rows are zero
    constexpr int NDB = 4;                              // accumulator buffers in TMEM (2 KPN columns each: main | corr)
    constexpr int A_BYTES = 16384;                      // 128 units x 32 reduction steps of FP32
    constexpr int B_BYTES = KPN * 128;                  // [KPN][32] floats
    constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
    constexpr uint32_t T_D0 = TF32_NAB * 64;            // first accumulator column
    static_assert(T_D0 + NDB * 2 * KPN <= 512, "tensor memory budget");
    pdl_prologue();
    if (*(volatile int *)&p.ctrl->stop) return;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *ring = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = (uint64_t *)(ring + (size_t)nstage * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + nstage;
    uint64_t *ta_full = empty_bar + nstage, *ta_empty = ta_full + TF32_NAB;
    uint64_t *d_full = ta_empty + TF32_NAB, *d_empty = d_full + 4;
    uint32_t *tmem_slot = (uint32_t *)(d_empty + 4);
    double *sGm = (double *)(tmem_slot + 2);            // KP*KP + 2 doubles: Gram side job
    int *gflag = (int *)(sGm + KP * KP + 2);
    const uint32_t ring_s = smem_u32(ring);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const long long u0 = (long long)blockIdx.x * p.U / G;
    const long long u1 = (long long)(blockIdx.x + 1) * p.U / G;
    const bool has_work = u0 < u1;
    const int tile0 = (int)(u0 / p.SPT), st0 = (int)(u0 - (long long)tile0 * p.SPT);
    const int nunits = (int)(u1 - u0);

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 5);                // 4 converter warps (A box read) + 1 MMA commit (B tiles read)
        }
        for (int s = 0; s < TF32_NAB; ++s) { mbar_init(&ta_full[s], 128); mbar_init(&ta_empty[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 12 && has_work) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp >= 9 && warp <= 11) {                      // Gram matrix of the fixed factor and its inverse (FP64, side job)
        gram_side_job<KB, WSTEP>(p, tid - 288, sGm, gflag);
        return;
    }
    if (!has_work) return;
    const uint32_t tbase = *(volatile uint32_t *)tmem_slot;

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0, tile = tile0, st = st0;
            uint32_t phase = 0;
            for (int it = 0; it < nunits; ++it) {
                mbar_wait(&empty_bar[stage], phase ^ 1, 1, it, stage);
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                unsigned char *dst = ring + (size_t)stage * STAGE_BYTES;
                if (!WSTEP) tma_load_2d(dst, &tmA, st * 32, tile * TILE_UNITS, &full_bar[stage]);
                else tma_load_3d(dst, &tmA, 0, st * 32, tile * 4, &full_bar[stage]);
                tma_load_2d(dst + A_BYTES, &tmBhi, st * 32, 0, &full_bar[stage]);
                tma_load_2d(dst + A_BYTES + B_BYTES, &tmBlo, st * 32, 0, &full_bar[stage]);
                if (++stage == nstage) { stage = 0; phase ^= 1; }
                if (++st == p.SPT) { st = 0; ++tile; }
            }
        }
        return;
    }

    if (warp >= 12 && warp < 16) {
        // ===== MMA issuers: warp 12 + q issues the stages it = q (mod 4) -- A buffer q, accumulator buffer it % NDB =====
        const int q = warp - 12;
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(128, KPN), idesc2 = umma_idesc_tf32(128, 2 * KPN);
            for (int it = q; it < nunits; it += 4) {
                const int stage = it % nstage;
                const uint32_t phase = (uint32_t)(it / nstage) & 1;
                const int ab = it & (TF32_NAB - 1), db = it % NDB;
                mbar_wait(&full_bar[stage], phase, 3, it, stage);                           // B tiles landed
                mbar_wait(&ta_full[ab], (uint32_t)(it / TF32_NAB) & 1, 4, it, ab);            // A operand in TMEM
                if (it >= NDB) mbar_wait(&d_empty[db], (uint32_t)(it / NDB - 1) & 1, 5, it, db);   // accumulators absorbed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t bhi = ring_s + (uint32_t)stage * STAGE_BYTES + A_BYTES;
                const uint32_t ta = tbase + 64 * ab, td = tbase + T_D0 + 2 * KPN * db;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dcat = umma_desc_sw128(bhi + 32 * ks);                    // rows 0..KPN-1 = B_hi, KPN..2KPN-1 = B_lo
                    umma_tf32_ts(td, ta + 8 * ks, dcat, idesc2, ks ? 1u : 0u);              // [main | corr] (+)= A_hi [B_hi ; B_lo]'
                    umma_tf32_ts(td + KPN, ta + 32 + 8 * ks, dcat, idesc, 1u);              // corr += A_lo B_hi'  (after the line above: same thread, in order)
                }
                umma_commit(&ta_empty[ab]);
                umma_commit(&empty_bar[stage]);
                umma_commit(&d_full[db]);
            }
        }
        __syncwarp();
        tf32_group_bar();                               // every converter has read its last accumulators
        if (warp == 12) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
        }
        return;
    }

    const bool timing = g_tf32_timing && blockIdx.x == 0 && (tid == 0 || tid == 512);
    const int agrp = (warp - 16) >> 2;                  // absorber group (warps >= 16)
    long long tmark = clock64();
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    if (warp >= 16) {
        // ===================== absorbers (warps 16-19 / 20-23): the accumulators of the even / odd stages -> registers ====
        // thread = TMEM lane = unit of the tile; the stage sums are added up as an unevaluated FP32 pair (two-sum:
        // ~48 bits, round-to-nearest) and widened to FP64 when the stream-K partial of a tile is written
        const int wq = warp & 3;
        const int unit = 32 * wq + lane;
        const uint32_t lanebase = (uint32_t)(32 * wq) << 16;
        const bool comp = g_tf32_comp != 0;
        float acc[KP], accl[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) acc[j] = accl[j] = 0.0f;
        int tile = tile0, st = st0;
        for (int it = 0; it < nunits; ++it) {
          if ((it & 1) == agrp) {
            const int db = it % NDB;
            TF32_TICK(5);
            mbar_wait(&d_full[db], (uint32_t)(it / NDB) & 1, 6, it, db);
            TF32_TICK(6);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t td = lanebase | (tbase + T_D0 + 2 * KPN * db);
            uint32_t vmain[KP], vcorr[KP];
#pragma unroll
            for (int c0 = 0; c0 < KP; c0 += 8) {
                tmem_ld8(td + c0, vmain + c0);
                tmem_ld8(td + KPN + c0, vcorr + c0);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&d_empty[db]);                       // the buffer goes back before the arithmetic
#pragma unroll
            for (int c0 = 0; c0 < KP; c0 += 8) {
                const uint32_t *vm = vmain + c0, *v1 = vcorr + c0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v = __fadd_rn(__uint_as_float(vm[j]), __uint_as_float(v1[j]));
                    const float a = acc[c0 + j];
                    const float sm = __fadd_rn(a, v);                       // two-sum: sm + e == a + v exactly
                    const float bp = __fsub_rn(sm, a);
                    const float e = __fadd_rn(__fsub_rn(a, __fsub_rn(sm, bp)), __fsub_rn(v, bp));
                    acc[c0 + j] = sm;
                    // The tensor core ADDS WITH TRUNCATION.  The main term is a sum of non-negative products
                    // (A >= 0 is checked at upload, the factors are NNLS solutions), so each of its 4 accumulation
                    // steps loses half a unit in the last place of the running sum on average: with running sums
                    // near v/4, v/2, 3v/4, v that is 1.125 or 1.375 ulp(v), depending on whether 3v/4 shares v's
                    // binade (measured: 1.38 +- 0.50 ulp, scripts/microbench/tf32_loss.cu).  The mean is put back;
                    // the scatter, and the signed errors of the correction terms, average out over the stages.
                    const uint32_t eb = vm[j] & 0x7f800000u;
                    const float ulp = (comp && eb > (24u << 23)) ? __uint_as_float(eb - (23u << 23)) : 0.0f;
                    const float cf = ((vm[j] & 0x007fffffu) >= 0x002aaaabu) ? 1.375f : 1.125f;
                    accl[c0 + j] = __fadd_rn(accl[c0 + j], __fmaf_rn(cf, ulp, e));
                }
            }
            TF32_TICK(7);
          }
            const bool last_of_tile = (st + 1 == p.SPT) || (it + 1 == nunits);
            if (last_of_tile) {
                // one partial per tile and absorber group (zeros if the group had no stage of this tile), like the
                // two consumer groups of the FP64 kernels: the solve kernel sums every slot
                const long long tb = (long long)tile * p.SPT;
                int slot;
                if (p.U >= G) slot = 2 * ((int)blockIdx.x - cta_of_unit(tb, p.U, G)) + agrp;
                else slot = 2 * (int)(u0 - tb) + agrp;
                double *Pb = p.P + ((size_t)tile * p.maxslots + slot) * (size_t)(KP * TILE_UNITS);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    Pb[(size_t)j * TILE_UNITS + unit] = (double)acc[j] + (double)accl[j];
                    acc[j] = accl[j] = 0.0f;
                }
            }
            if (++st == p.SPT) { st = 0; ++tile; }
        }
        if (timing)
            for (int q = 5; q < 8; ++q) g_tf32_ticks[q] += tacc[q];
        tf32_group_bar();
        return;
    }

    // ===================== converters: two groups of four warps, stage it belongs to group it & 1 =====================
    const int grp = warp >> 2, wq = warp & 3;
    const int unit = 32 * wq + lane;                    // TMEM lane = unit of the tile (column of A / row of A)
    const uint32_t lanebase = (uint32_t)(32 * wq) << 16;
    int stage = grp;
    uint32_t phase = 0;
    for (int it = grp; it < nunits; it += 2) {
        const uint32_t sb = ring_s + (uint32_t)stage * STAGE_BYTES;
        TF32_TICK(0);
        mbar_wait(&full_bar[stage], phase, 2, it, stage);
        TF32_TICK(1);
        const int ab = it & (TF32_NAB - 1);
        // two halves of 16 reduction indices (register budget: 768 threads): load, split, store to tensor memory
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t hi[16], lo[16];
            if (!WSTEP) {
                // box [128 columns][32 rows of floats] (128-byte rows, 16-byte chunk q at q ^ (column & 7)): this
                // thread's column, four 16-byte loads per half
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = lds128f(sb + (uint32_t)unit * 128 + (uint32_t)(((4 * hf + q) ^ (unit & 7)) << 4));
                    hi[4 * q + 0] = __float_as_uint(v.x); hi[4 * q + 1] = __float_as_uint(v.y);
                    hi[4 * q + 2] = __float_as_uint(v.z); hi[4 * q + 3] = __float_as_uint(v.w);
                }
            } else {
                // box [4 row groups][32 columns][32 rows of floats]: row group = this warp's quarter, row = lane;
                // one 4-byte load per column (a warp reads one whole 128-byte row: conflict-free)
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    hi[c] = lds32(sb + (uint32_t)wq * 4096 + (uint32_t)(16 * hf + c) * 128 +
                                  (uint32_t)((((lane >> 2) ^ (c & 7)) << 4) | ((lane & 3) << 2)));
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float x = __uint_as_float(hi[c]);
                const uint32_t h = (hi[c] + 0x1000u) & 0xffffe000u;    // nearest TF32 number
                hi[c] = h;
                lo[c] = __float_as_uint(x - __uint_as_float(h));       // signed remainder, exact in FP32
            }
            if (hf == 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);        // the A box of this stage has been read
            }
            if (hf == 0) {
                TF32_TICK(2);
                if (it >= TF32_NAB) mbar_wait(&ta_empty[ab], (uint32_t)(it / TF32_NAB - 1) & 1, 7, it, ab);
                TF32_TICK(3);
            }
            tmem_st16(lanebase | (tbase + 64 * ab + 16 * hf), hi);
            tmem_st16(lanebase | (tbase + 64 * ab + 32 + 16 * hf), lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&ta_full[ab]);
        TF32_TICK(4);
        stage += 2;
        if (stage >= nstage) { stage -= nstage; phase ^= 1; }
    }
    if (timing)
        for (int q = 0; q < 5; ++q) g_tf32_ticks[q] += tacc[q];
    tf32_group_bar();
}

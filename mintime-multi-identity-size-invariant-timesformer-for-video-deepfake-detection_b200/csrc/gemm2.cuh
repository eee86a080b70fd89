// 2-CTA (cta_group::2) variant of the tcgen05 GEMM for the large transformer contractions.
// Included by gemm.cu inside its anonymous namespace (uses PipeState, staging_piece, fast_gelu, ...).
//
// Why: with one CTA per 128x256 tile every k-block pulls 48 KiB of operands from L2 per 512 MMA cycles
// (94 B/clk/SM), but an SM can only pull ~42 B/clk from L2 -- the round-1 kernels sat at 45-50 % tensor-pipe
// for exactly that reason (QKV 43 us, FF2 61 us match bytes / 42 B/clk).  A CTA pair computes a 256x256 tile:
// each CTA loads its own 128 rows of A and HALF of the W tile (128 rows), the pair's MMA reads both halves,
// so per-SM operand traffic drops to 32 KiB per k-block (1.5x less) and 6 pipeline stages fit.
//
//   warp 0      TMA producer (both CTAs; bytes are credited to the leader CTA's full barrier)
//   warp 1      MMA issuer (leader CTA only): tcgen05.mma.cta_group::2, M = 256, N = 256; accumulator rows
//               0-127 live in the leader's TMEM, rows 128-255 in the peer's; double buffered (2 x 256 cols)
//   warps 2..   epilogue (4 or 8 warps per CTA): own TMEM rows -> bias / swish / GEGLU / fp32 reduce-add ->
//               swizzled staging -> TMA store; 8 warps (two groups alternating 128-byte column chunks) for
//               the math-heavy GEGLU epilogue
#pragma once

constexpr int kStage2Bytes = 2 * kAStageBytes;   // A 128x64 + this CTA's half of W (128x64)
// pipeline depth: 6 x 32 KiB + one epilogue group's staging (32 KiB), or 5 x 32 KiB + two groups' (64 KiB)
constexpr int stages2(int epi_warps) { return epi_warps == 8 ? 5 : 6; }

struct Tc2Params {
  int M, N, K;
  int tiles_m, tiles_n;    // 256 x 256 pair tiles
  EpiParams epi;
};

// Epilogue of ONE accumulator tile of this CTA (128 rows x 256 TMEM columns at `taddr`): bias / swish / GEGLU / fp32
// reduce-add -> swizzled staging -> TMA store.  `grp` = this warp's group of 4 (groups alternate the 128-byte column
// chunks), STG_BUFS staging buffers per group; the leader's `tmem_empty_bar` gets one arrival per epilogue thread once the
// thread's last TMEM read of the tile has completed.
template <int KIND, int EPI_WARPS, int STG_BUFS>
__device__ __forceinline__ void tc2_epilogue_tile(const CUtensorMap* tmap_out, const EpiParams& e, uint32_t taddr, int m0, int n0,
                                                  int grp, int trow, bool issuer, uint8_t* my_stage, uint32_t& store_it,
                                                  uint64_t* tmem_empty_bar) {
  constexpr int kGroups = EPI_WARPS / 4;
  constexpr int kChunk = KIND == EPI_GEGLU ? 128 : (KIND == EPI_RESID_F32 ? 32 : 64);
  constexpr int kChunks = 256 / kChunk;
  for (int ci = grp; ci < kChunks; ci += kGroups) {
    const int c = ci * kChunk;
    uint8_t* stg = my_stage + (store_it % STG_BUFS) * kStagingBytes;
    if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STG_BUFS - 1) : "memory");
    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    if (KIND == EPI_GEGLU) {
#pragma unroll
      for (int ob = 0; ob < 2; ++ob) {
        uint32_t ru[32], rg[32];
        ptx::tmem_ld_32x32b_x32(taddr + c + ob * 64, ru);
        ptx::tmem_ld_32x32b_x32(taddr + c + ob * 64 + 32, rg);
        ptx::tmem_ld_wait();
        const int col = n0 + c + ob * 64;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float2 u2[4], g2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            u2[i] = make_float2(__uint_as_float(ru[g * 8 + 2 * i]), __uint_as_float(ru[g * 8 + 2 * i + 1]));
            g2[i] = make_float2(__uint_as_float(rg[g * 8 + 2 * i]), __uint_as_float(rg[g * 8 + 2 * i + 1]));
          }
          if (e.bias) {
            float b[8];
            load8(e.bias + col + g * 8, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) u2[i] = __fadd2_rn(u2[i], make_float2(b[2 * i], b[2 * i + 1]));
            load8(e.bias + col + 32 + g * 8, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) g2[i] = __fadd2_rn(g2[i], make_float2(b[2 * i], b[2 * i + 1]));
          }
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 v = geglu2(u2[i], g2[i]);
            o[i] = pack_bf16(v.x, v.y);
          }
          *staging_piece(stg, trow, ob * 4 + g) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    } else {
      constexpr int kQ = kChunk / 16;
      uint32_t r[kQ][16];
#pragma unroll
      for (int q = 0; q < kQ; ++q) ptx::tmem_ld_32x32b_x16(taddr + c + q * 16, r[q]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int col = n0 + c + q * 16 + hh * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[q][hh * 8 + i]);
          float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (e.bias) load8(e.bias + col, b);
          if (KIND == EPI_STORE && e.act == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = swish_shift_fast(v[i], b[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += b[i];
          }
          if (KIND == EPI_STORE) {
            *staging_piece(stg, trow, q * 2 + hh) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          } else {
            *staging_piece(stg, trow, q * 4 + hh * 2) =
                make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
            *staging_piece(stg, trow, q * 4 + hh * 2 + 1) =
                make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
          }
        }
      }
    }
    if (ci + kGroups >= kChunks) {             // this group's last TMEM read of the tile
      ptx::tc_fence_before();
      ptx::mbar_arrive_cluster(tmem_empty_bar, 0);   // the leader's MMA warp waits for both CTAs
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    if (issuer) {
      const int ocol = KIND == EPI_GEGLU ? (n0 + c) / 2 : n0 + c;
      if (KIND == EPI_RESID_F32)
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tmap_out)), "r"(ptx::smem_u32(stg)), "r"(ocol), "r"(m0)
                     : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tmap_out)), "r"(ptx::smem_u32(stg)), "r"(ocol), "r"(m0)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    ++store_it;
  }
}

template <int KIND, int EPI_WARPS>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_out, Tc2Params p) {
  constexpr int kGroups = EPI_WARPS / 4;           // epilogue warp groups (each owns 2 staging buffers)
  constexpr int kStages2 = stages2(EPI_WARPS);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages2 * kAStageBytes;
  uint8_t* smem_stage = smem_b + kStages2 * kAStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + kGroups * 2 * kStagingBytes);
  uint64_t* full_bar = bars;                       // used in the leader CTA
  uint64_t* empty_bar = bars + kStages2;           // both CTAs (multicast commit)
  uint64_t* tmem_full = bars + 2 * kStages2;       // both CTAs (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;            // used in the leader CTA (both epilogues arrive)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();    // 0 = leader
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < kStages2; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 2 * EPI_WARPS * 32);
    }
    ptx::fence_mbar_init();
  }
  ptx::cluster_sync_all();                         // barriers of both CTAs initialised before any remote use
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_ptr_smem, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      PipeState ps;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / p.tiles_n) * 256 + (int)rank * 128;
        const int n0 = (tile % p.tiles_n) * 256 + (int)rank * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[ps.stage], 2u * kStage2Bytes);   // both CTAs' bytes
          ptx::tma_load_2d_2sm(smem_a + ps.stage * kAStageBytes, &tmap_a, &full_bar[ps.stage], kb * kBlockK, m0);
          ptx::tma_load_2d_2sm(smem_b + ps.stage * kAStageBytes, &tmap_b, &full_bar[ps.stage], kb * kBlockK, n0);
          ps.advance(kStages2);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA)
    if (leader && lane == 0) {
      PipeState ps;
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(256, 256);
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[ps.stage], ps.phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + ps.stage * kAStageBytes);
          const uint32_t b_addr = ptx::smem_u32(smem_b + ps.stage * kAStageBytes);
          const int k_left = p.K - kb * kBlockK;
          const int ksteps = k_left >= kBlockK ? 4 : (k_left + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            ptx::umma_bf16_ss_2sm(tmem_d, ptx::umma_smem_desc_sw128(a_addr + k * 32),
                                  ptx::umma_smem_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit_2sm(&empty_bar[ps.stage], 0x3);   // frees the stage in BOTH CTAs
          ps.advance(kStages2);
        }
        ptx::umma_commit_2sm(&tmem_full[as], 0x3);            // accumulator ready in BOTH CTAs
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (EPI_WARPS warps)
    const int ew = warp - 2;
    const int grp = ew >> 2;                       // warp group: alternates 128-byte column chunks
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    const int trow = quad * 32 + lane;
    const bool issuer = (ew & 3) == 0 && lane == 0;
    const EpiParams& e = p.epi;
    uint8_t* my_stage = smem_stage + grp * 2 * kStagingBytes;
    uint32_t store_it = 0;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / p.tiles_n) * 256 + (int)rank * 128;
      const int n0 = (tile % p.tiles_n) * 256;
      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256);
      tc2_epilogue_tile<KIND, EPI_WARPS, 2>(&tmap_out, e, taddr, m0, n0, grp, trow, issuer, my_stage, store_it, &tmem_empty[as]);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();                         // nobody may still target this CTA's smem / TMEM
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// A-RESIDENT variant for K = 512 (the transformer's FF1 / GEGLU and to_qkv contractions).
// Measured in round 2 (ncu, profiles/r2_fused_attn_gemm_ncu_summary.txt): the streaming kernel above re-reads A once per
// column tile and W once per row tile -- 822 MB of L2 -> SM traffic per GEGLU launch, 8 TB/s over the 100 us it takes --
// and its MMA warp spends ~40 % of the time waiting for operands while the epilogue warps wait for the MMA: the kernel
// is bound by the L2 -> SM fabric, not by the tensor pipe (one K = 16 tcgen05.mma per ~185 clk = 66 us of issue).
// Here a CTA keeps its 128 x 512 A tile (8 k-blocks, 128 KiB) in shared memory for ALL column tiles of the row tile
// and only W streams (4- or 5-stage ring): A is read from L2 once, the traffic drops to 437 MB.  The linearised
// (row tile, column tile) steps are split evenly over the CTA pairs, as in attention_fused.cuh.
//   warp 0: W producer   warp 1: MMA issuer (leader) + TMEM   warp 2: A producer   warps 3..: epilogue (tc2_epilogue_tile)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kResKB = 8;                        // K = 512
// W ring depth: what fits beside the 128 KiB A tile and the staging buffers (8 epilogue warps: 2 x 16 KiB -> 4 stages)
constexpr int res_b_stages(int epi_warps) { return epi_warps == 8 ? 4 : 5; }

template <int KIND, int EPI_WARPS>
__global__ void __launch_bounds__(96 + 32 * EPI_WARPS, 1)
gemm_tc2a_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, Tc2Params p) {
  constexpr int kGroups = EPI_WARPS / 4;
  constexpr int kResBStages = res_b_stages(EPI_WARPS);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* smem_a = smem;                                          // [kResKB][128 x 64]
  uint8_t* smem_b = smem_a + kResKB * kAStageBytes;                // [kResBStages][128 x 64]
  uint8_t* smem_stage = smem_b + kResBStages * kAStageBytes;       // [kGroups][128 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + kGroups * kStagingBytes);
  uint64_t* a_full = bars;                         // [kResKB] leader
  uint64_t* a_empty = a_full + kResKB;             // [kResKB] both (multicast commit)
  uint64_t* b_full = a_empty + kResKB;             // [kResBStages] leader
  uint64_t* b_empty = b_full + kResBStages;        // both
  uint64_t* tmem_full = b_empty + kResBStages;     // [2] both
  uint64_t* tmem_empty = tmem_full + 2;            // [2] leader
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    ptx::prefetch_tmap(&tmap_out);
    for (int i = 0; i < kResKB; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kResBStages; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tmem_full[i], 1); ptx::mbar_init(&tmem_empty[i], 2 * EPI_WARPS * 32); }
    ptx::fence_mbar_init();
  }
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_ptr_smem, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total = p.tiles_m * p.tiles_n;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int s0 = (int)((long long)total * pair / num_pairs), s1 = (int)((long long)total * (pair + 1) / num_pairs);

  if (warp == 2) {
    // ===================================================================== A producer
    if (lane == 0) {
      int gen = 0;
      for (int step = s0; step < s1; ++gen) {
        const int tm = step / p.tiles_n, tn = step - tm * p.tiles_n;
        const int m0 = tm * 256 + (int)rank * 128;
        for (int kb = 0; kb < kResKB; ++kb) {
          ptx::mbar_wait(&a_empty[kb], (gen & 1) ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&a_full[kb], 2u * kAStageBytes);
          ptx::tma_load_2d_2sm(smem_a + kb * kAStageBytes, &tmap_a, &a_full[kb], kb * kBlockK, m0);
        }
        step += min(p.tiles_n - tn, s1 - step);
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===================================================================== W producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = s0; step < s1; ++step) {
        const int n0 = (step % p.tiles_n) * 256 + (int)rank * 128;
        for (int kb = 0; kb < kResKB; ++kb) {
          ptx::mbar_wait(&b_empty[stage], phase ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&b_full[stage], 2u * kAStageBytes);
          ptx::tma_load_2d_2sm(smem_b + stage * kAStageBytes, &tmap_b, &b_full[stage], kb * kBlockK, n0);
          if (++stage == kResBStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA)
    if (leader && lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(256, 256);
      int stage = 0, gen = 0, it = 0;
      uint32_t phase = 0;
      for (int step = s0; step < s1; ++step, ++it) {
        const int tn = step % p.tiles_n;
        const bool first = step == s0 || tn == 0, last = tn == p.tiles_n - 1 || step == s1 - 1;
        const int as = it & 1;
        ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
        for (int kb = 0; kb < kResKB; ++kb) {
          if (first) ptx::mbar_wait(&a_full[kb], gen & 1);
          ptx::mbar_wait(&b_full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + kb * kAStageBytes);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * kAStageBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss_2sm(tmem_d, ptx::umma_smem_desc_sw128(a_addr + k * 32), ptx::umma_smem_desc_sw128(b_addr + k * 32),
                                  idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit_2sm(&b_empty[stage], 0x3);
          if (last) ptx::umma_commit_2sm(&a_empty[kb], 0x3);
          if (++stage == kResBStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_2sm(&tmem_full[as], 0x3);
        if (last) ++gen;
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue
    const int ew = warp - 3;
    const int grp = ew >> 2;
    const int quad = warp & 3;
    const int trow = quad * 32 + lane;
    const bool issuer = (ew & 3) == 0 && lane == 0;
    uint8_t* my_stage = smem_stage + grp * kStagingBytes;
    uint32_t store_it = 0;
    int it = 0;
    for (int step = s0; step < s1; ++step, ++it) {
      const int as = it & 1;
      const int tm = step / p.tiles_n, tn = step - tm * p.tiles_n;
      const int m0 = tm * 256 + (int)rank * 128, n0 = tn * 256;
      ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256);
      tc2_epilogue_tile<KIND, EPI_WARPS, 1>(&tmap_out, p.epi, taddr, m0, n0, grp, trow, issuer, my_stage, store_it, &tmem_empty[as]);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int KIND, int EPI_WARPS>
int launch_tc2a(const GemmArgs& g, cudaStream_t stream) {
  Tc2Params p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.epi = g.epi;
  p.tiles_m = (g.M + 255) / 256;
  p.tiles_n = g.N / 256;
  constexpr int kGroups = EPI_WARPS / 4;
  constexpr int kResBStages = res_b_stages(EPI_WARPS);
  const size_t smem = (size_t)(kResKB + kResBStages) * kAStageBytes + (size_t)kGroups * kStagingBytes + 1024 +
                      (2 * kResKB + 2 * kResBStages + 4) * 8 + 16;
  CUtensorMap ta, tb, tout;
  int rc = make_tmap_2d(&ta, g.a, false, g.M, g.K, g.K, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, g.w, false, g.N, g.K, g.K, 128);
  if (rc) return rc;
  const bool f32 = KIND == EPI_RESID_F32;
  rc = make_tmap_2d(&tout, g.epi.out, f32, g.M, KIND == EPI_GEGLU ? g.N / 2 : g.N, g.epi.ldo, 128);
  if (rc) return rc;
  auto kern = gemm_tc2a_kernel<KIND, EPI_WARPS>;
  if (first_use_on_device(reinterpret_cast<const void*>(kern))) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(gemm_tc2a)");
  }
  const int pairs = std::min(p.tiles_m * p.tiles_n, num_sms() / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(96 + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, p);
  if (e != cudaSuccess) return cuda_status(e, "cudaLaunchKernelEx(gemm_tc2a)");
  MT_LAUNCH_CHECK("gemm_tc2a_kernel");
  return MT_OK;
}

// the A-resident schedule pays where A is re-read many times: K = 512 exactly and at least 4 column tiles
inline bool tc2a_eligible(const GemmArgs& g) { return g.K == kResKB * kBlockK && g.N >= 1024; }

template <int KIND, int EPI_WARPS>
int launch_tc2(const GemmArgs& g, cudaStream_t stream) {
  Tc2Params p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.epi = g.epi;
  p.tiles_m = (g.M + 255) / 256;
  p.tiles_n = g.N / 256;
  constexpr int kGroups = EPI_WARPS / 4;
  constexpr int kStages2 = stages2(EPI_WARPS);
  const size_t smem = (size_t)kStages2 * kStage2Bytes + (size_t)kGroups * 2 * kStagingBytes + 1024 +
                      (2 * kStages2 + 4) * 8 + 16;
  CUtensorMap ta, tb, tout;
  int rc = make_tmap_2d(&ta, g.a, false, g.M, g.K, g.K, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, g.w, false, g.N, g.K, g.K, 128);
  if (rc) return rc;
  const bool f32 = KIND == EPI_RESID_F32;
  rc = make_tmap_2d(&tout, g.epi.out, f32, g.M, KIND == EPI_GEGLU ? g.N / 2 : g.N, g.epi.ldo, 128);
  if (rc) return rc;
  auto kern = gemm_tc2_kernel<KIND, EPI_WARPS>;
  if (first_use_on_device(reinterpret_cast<const void*>(kern))) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(gemm_tc2)");
  }
  int pairs = std::min(p.tiles_m * p.tiles_n, num_sms() / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(64 + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, p);
  if (e != cudaSuccess) return cuda_status(e, "cudaLaunchKernelEx(gemm_tc2)");
  MT_LAUNCH_CHECK("gemm_tc2_kernel");
  return MT_OK;
}

// eligibility: plain store / GEGLU / fp32 residual epilogues, N a multiple of the 256-wide pair tile,
// enough rows to fill the chip, 16-byte aligned output rows
inline bool tc2_eligible(const GemmArgs& g) {
  if (g.gate || g.epi.resid) return false;
  if (g.epi.kind != EPI_STORE && g.epi.kind != EPI_GEGLU && g.epi.kind != EPI_RESID_F32) return false;
  // (K < 384: the pair kernel's longer prologue / epilogue per tile does not pay: the 320 -> 1280 head conv runs
  // 43 us on the single-CTA kernel against 70 us here)
  if (g.N % 256 != 0 || g.M < 4096 || g.K < 384) return false;
  if ((g.epi.ldo * (g.epi.kind == EPI_RESID_F32 ? 4 : 2)) % 16 != 0 || (reinterpret_cast<uintptr_t>(g.epi.out) & 15)) return false;
  return true;
}

// syrk.cu -- K3: kinship K = X X^T on tcgen05 tensor cores (sm_100a).
//
// Replaces the block loop of SnpReader._read_kernel (pysnptools/snpreader/snpreader.py:651-655: read +
// standardize + val.dot(val.T) + `K +=`; GEMM call site snpdata.py:203-206).
//
// Per chunk of SNPs:
//   1. per-SNP statistics from exact 2-bit counts (decode.cu, no output matrix);
//   2. k_absmax: largest |standardized value| of the chunk -> one power-of-two scale so fp16 never overflows
//      or loses small SNPs to subnormals;
//   3. k_planes: decode + standardize + split every value x*scale into fp16 hi + fp16 lo and store the two
//      operand planes [n_pad, k_pad] K-major (SNP index fastest) -- the only time X exists, as 4 bytes/genotype
//      of scratch that is consumed from L2/HBM by TMA;
//   4. k_syrk: persistent warp-specialised tcgen05 kernel over the lower-triangular 128x256 tiles:
//      TMA (128B swizzle) -> 2-stage smem ring -> three MMAs per k-step (hi*hi + hi*lo + lo*hi; the dropped
//      lo*lo term is 2^-22 relative) accumulating fp32 in tensor memory.  The tensor core truncates (rounds
//      toward zero) on every accumulate, which biases long positive sums (the diagonal of K) by about
//      3e-8 per MMA; so TMEM only ever holds a short run (RUN_KB k-blocks), two 256-column accumulators
//      alternate, and eight epilogue warps drain each finished run with tcgen05.ld into register-resident
//      fp32 sums (round-to-nearest adds) while the next run is being multiplied.  One write of
//      K (+)= sum / scale^2 per tile and chunk.
//   5. k_mirror copies the lower triangle into the upper one.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <atomic>
#include <cuda_fp8.h>
#include "pstb_common.cuh"

namespace pstb {
int read_impl_ex(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                 int count_a1, int mode, double a, double b, int use_stats, double* d_stats, void* d_out, int dtype, int order,
                 void* stream, unsigned int* d_miss_flag);

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;                      // 16 KiB  (one fp16 plane tile)
constexpr int B_BYTES = BN * BK * 2;                      // 32 KiB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;    // hi + lo of both operands: 96 KiB
constexpr int SYRK_SMEM = STAGES * STAGE_BYTES + 1024;    // + slack for the 1024-byte alignment of swizzled tiles
constexpr int SYRK_THREADS = 320;                         // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int RUN_KB = 4;                                 // k-blocks accumulated in TMEM before a drain (48 MMAs)
constexpr int TMEM_COLS = 512;                            // two 128x256 fp32 accumulators
constexpr int ROW_PAD = 256;                              // plane rows padded so every TMA box is in bounds
constexpr int GROUP_I = 16, GROUP_J = 8;                  // tile rasterisation: 2048 x 2048 super-blocks stay L2 resident

// scalars block in the workspace
struct Scalars {
    unsigned int absmax_bits;    // float bits of max |x| over the chunk (atomicMax on positive floats): 3-term split of x
    unsigned int absmax_w_bits;  // float bits of max |w (g - mean)| over the chunk: exact-dosage path
    unsigned int need_3term;     // != 0: the chunk cannot take the exact-dosage path (a trained mean outside [0, 2], or forced): 3-term split
    unsigned int pad[61];
};

// Exact-dosage path.  x = m (g - mu) f with m = 1 for an observed genotype, 0 for a missing one (mean imputation,
// standardizer.py:145-163), f = 1/sd (Unit) or BetaPDF (Beta), w = f^2:   K_ik = sum_j T_ij w_j T_kj,  T = m (g - mu).
// Round mu to 10 fractional bits, mu' = mu + delta.  The LEFT plane holds
//     L_ij = g_ij - mu'_j      (observed)   -- exact in fp16: |L| <= 2 with 10 fractional bits
//          = fp16(-delta_j)    (missing)    -- |delta| <= 2^-11, so fp16 carries it to 2^-22
// so that T_ij = L_ij + delta_j for EVERY entry (missing ones to 2^-22), whatever the missing pattern.  The RIGHT plane holds
// the exact weighted value B_kj = w_j T_kj (0 for a missing genotype), split hi + lo.  Then
//     K_ik = sum_j L_ij B_kj  +  v_k,      v_k = sum_j delta_j B_kj   (rank one, accumulated in fp64 by k_planes)
// : 2 MMAs per k-step instead of 3, no dropped term, for data WITH missing genotypes and for trained statistics as well (the
// round-1 kernel fell back to the 3-term split whenever a chunk had one missing code).  Because L is centred the products have
// the magnitude of K itself, so the truncating accumulator is as harmless as in the 3-term path.
// Low term on the fp8 pipe: e4m3 carries 4 significant bits, so with mu'' = mu' rounded to a multiple of 1/8 the left factor
// g - mu'' is EXACT in e4m3 as well; L = (g - mu'') + d, d = mu'' - mu' (|d| <= 1/16), and the d part is rank one again:
//     sum_j L_ij lo_kj = sum_j (g_ij - mu''_j) lo8_kj + sum_j d_j lo8_kj       (only lo is rounded: sqrt(2) less error than
// rounding both factors; CPU emulation scripts/emulate_masked_split.py).
__device__ __forceinline__ double round_mu(double mean) { return rint(mean * 1024.0) * (1.0 / 1024.0); }
__device__ __forceinline__ double weight_of(int mode, double mean, double sd, double a, double b, double lnB) {
    if (!(sd == sd) || isinf(sd) || !(mean == mean)) return 0.0;
    const double f = (mode == PSTB_STD_BETA) ? beta_factor(mean, a, b, lnB) : 1.0 / sd;
    return f * f;
}

__device__ __forceinline__ int scale_exponent(unsigned int absmax_bits) {
    // scale = 2^(14 - e) with absmax < 2^e, clamped so 1/scale^2 stays a normal float
    float m = __uint_as_float(absmax_bits);
    int e = 14;
    if (m > 0.0f && m < INFINITY) (void)frexpf(m, &e);
    e = e < -45 ? -45 : (e > 60 ? 60 : e);
    return 14 - e;
}

__global__ void k_absmax(const double* stats, long long ns, int mode, double a, double b, double lnB, Scalars* sc) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f, mw = 0.0f;
    if (s < ns) {
        const double mean = stats[2 * s], sd = stats[2 * s + 1];
        const double f = (mode == PSTB_STD_BETA) ? beta_factor(mean, a, b, lnB) : 0.0;
        for (int g = 0; g < 3; ++g) {
            double v = fabs(std_value(mode, (double)g, mean, sd, f));
            if (v == v && v < 1e300) m = fmaxf(m, __double2float_ru(v));
        }
        const double w = weight_of(mode, mean, sd, a, b, lnB);
        const double hv = w * fmax(fabs(mean), fabs(2.0 - mean)) * (1.0 + 1e-6);
        if (hv == hv && hv < 1e300) mw = fmaxf(mw, __double2float_ru(hv));
        // a (trained) mean outside [0, 2]: g - mu' would need more than 11 significant bits
        if (w > 0.0 && !(mean >= 0.0 && mean <= 2.0)) sc->need_3term = 1u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    }
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(&sc->absmax_bits, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && mw > 0.0f) atomicMax(&sc->absmax_w_bits, __float_as_uint(mw));
}

// ---- operand planes --------------------------------------------------------------------------------------
constexpr int PT_S = 64, PT_I = 256, PT_PITCH = PT_I / 4 + 2;   // 66-byte pitch: conflict-free for 2 SNPs per lane

struct PlaneParams {
    const uint8_t* packed;
    long long ld, iid_count, sid_count;
    Axis iid, sid;          // sid already offset to the chunk
    int count_a1, mode;
    double a, b, lnB;
    const double* stats;    // [ns][2] of the chunk
    const Scalars* sc;
    __half* hi;             // plane 0: x_hi   (exact-dosage path: h)
    __half* lo;             // plane 1: x_lo   (exact-dosage path: (w h)_hi)
    __half* p2;             // plane 2: unused (exact-dosage path: B_lo; with fp8lo: two byte planes, B_lo then g - mu'', e4m3)
    long long p2_row0, p2_rows;   // rows of plane 2 before this operand / in total (0: this operand alone, n_pad rows)
    int fp8lo;              // exact-dosage path with the low term on the fp8 pipe (see k_syrk2)
    double* u;              // [n_pad] rank-one vector v of the exact-dosage path, accumulated over chunks (NULL: not wanted)
    long long n_pad, k_pad;
    int dense;
    long long byte_off;
};

template <int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) k_planes(const PlaneParams p) {
    __shared__ __align__(4) unsigned char codes[PT_S][PT_PITCH];
    __shared__ __align__(8) __half lut_hi[PT_S][4], lut_lo[PT_S][4];
    const long long tiles_i = p.n_pad / PT_I;
    const long long ts = blockIdx.x / tiles_i, ti = blockIdx.x % tiles_i;
    const long long b0 = ts * PT_S, i0 = ti * PT_I;
    const long long n_out = p.iid.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // the tile's 2-bit codes: 64 SNPs x 64 bytes.  One 32-bit load per 16 individuals where the rows are dense and word-aligned (the
    // byte loads this replaces left every CTA waiting on 16 dependent-latency loads per thread: stall long_scoreboard 5.3 per issue)
    const bool word_ok = p.dense && ((reinterpret_cast<uintptr_t>(p.packed) + (uintptr_t)p.byte_off) & 3u) == 0 && (p.ld & 3) == 0;
    for (int e = threadIdx.x; e < PT_S * (PT_I / 16); e += blockDim.x) {
        const int s = e / (PT_I / 16), w = e % (PT_I / 16);
        uint32_t word = 0x55555555u;                           // padding decodes as "missing" -> 0
        if (b0 + s < p.sid.n && i0 + 16 * w < n_out) {
            long long j = p.sid.at(b0 + s);
            j = j < 0 ? 0 : (j >= p.sid_count ? p.sid_count - 1 : j);
            const uint8_t* src = p.packed + j * p.ld;
            if (word_ok && i0 + 16 * w + 15 < n_out) {
                word = __ldg(reinterpret_cast<const uint32_t*>(src + p.byte_off + (i0 >> 2)) + w);
            } else {
                word = 0;
                for (int qq = 0; qq < 4; ++qq) {
                    const int q = 4 * w + qq;
                    uint32_t byte = 0x55u;
                    if (i0 + 4 * q < n_out) {
                        if (p.dense && i0 + 4 * q + 3 < n_out) {
                            byte = __ldg(src + p.byte_off + (i0 >> 2) + q);
                        } else {
                            byte = 0;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                long long a = i0 + 4 * q + t;
                                uint32_t code = 1u;
                                if (a < n_out) {
                                    long long i = p.iid.at(a);
                                    i = i < 0 ? 0 : (i >= p.iid_count ? p.iid_count - 1 : i);
                                    code = ((uint32_t)__ldg(src + (i >> 2)) >> (2 * (i & 3))) & 3u;
                                }
                                byte |= code << (2 * t);
                            }
                        }
                    }
                    word |= byte << (8 * qq);
                }
            }
        }
        // rows of `codes` are 66 bytes apart: 2-byte aligned
        *reinterpret_cast<uint16_t*>(&codes[s][4 * w]) = (uint16_t)word;
        *reinterpret_cast<uint16_t*>(&codes[s][4 * w + 2]) = (uint16_t)(word >> 16);
    }
    const bool fast = p.sc->need_3term == 0u;
    __shared__ __align__(8) __half lut_p2[PT_S][4];
    __shared__ __align__(4) unsigned char lut_l8[PT_S][4], lut_h8[PT_S][4];   // fp8 (e4m3) low term and left factor
    __shared__ float lut_u[PT_S][4];
    if (threadIdx.x < PT_S) {
        const int s = threadIdx.x;
        double v[4] = {0.0, 0.0, 0.0, 0.0};                     // right operand by 2-bit code (code 1 = missing stays 0)
        double hv[4] = {0.0, 0.0, 0.0, 0.0};                    // exact-dosage path: left operand L
        double h8v[4] = {0.0, 0.0, 0.0, 0.0};                   // ... and its e4m3-exact twin g - mu''
        double delta = 0.0, d8 = 0.0, inv_scale = 0.0;
        if (b0 + s < p.sid.n) {
            const double mean = p.stats[2 * (b0 + s)], sd = p.stats[2 * (b0 + s) + 1];
            if (!fast) {
                const double f = (p.mode == PSTB_STD_BETA) ? beta_factor(mean, p.a, p.b, p.lnB) : 0.0;
                const double scale = ldexp(1.0, scale_exponent(p.sc->absmax_bits));
                const double v0 = std_value(p.mode, 0.0, mean, sd, f) * scale, v1 = std_value(p.mode, 1.0, mean, sd, f) * scale,
                             v2 = std_value(p.mode, 2.0, mean, sd, f) * scale;
                v[0] = p.count_a1 ? v2 : v0;
                v[2] = v1;
                v[3] = p.count_a1 ? v0 : v2;
            } else {
                const double w = weight_of(p.mode, mean, sd, p.a, p.b, p.lnB);
                if (w > 0.0 && w < 1e300) {                     // SNC / all-missing / NaN statistics: the SNP contributes nothing
                    const double mu = round_mu(mean), mu8 = rint(mu * 8.0) * 0.125;
                    const int se = scale_exponent(p.sc->absmax_w_bits);
                    const double scale = ldexp(1.0, se);
                    inv_scale = ldexp(1.0, -se);
                    delta = mu - mean;
                    d8 = mu8 - mu;
                    const double g0 = p.count_a1 ? 2.0 : 0.0, g3 = p.count_a1 ? 0.0 : 2.0;
                    hv[0] = g0 - mu;  hv[2] = 1.0 - mu;  hv[3] = g3 - mu;
                    hv[1] = (double)__half2float(__float2half_rn((float)(-delta)));          // missing: T = L + delta = 0 to 2^-22
                    h8v[0] = g0 - mu8; h8v[2] = 1.0 - mu8; h8v[3] = g3 - mu8;
                    h8v[1] = -d8;                                                             // missing: (g - mu'') + d ~ 0
                    v[0] = w * (g0 - mean) * scale; v[2] = w * (1.0 - mean) * scale; v[3] = w * (g3 - mean) * scale;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double x = v[c];
            if (!(x == x) || fabs(x) > 60000.0) x = 0.0;        // NaN statistics (all-missing SNP) contribute nothing
            const __half h = __float2half_rn((float)x);
            const double res = x - (double)__half2float(h);
            const __half l = __float2half_rn((float)res);
            if (!fast) {
                lut_hi[s][c] = h;
                lut_lo[s][c] = l;
                lut_p2[s][c] = __float2half_rn(0.0f);
                lut_l8[s][c] = lut_h8[s][c] = 0;
                lut_u[s][c] = 0.0f;
            } else {
                lut_hi[s][c] = __float2half_rn((float)hv[c]);   // exact: |L| <= 2 with 10 fractional bits
                lut_lo[s][c] = h;
                lut_p2[s][c] = l;
                // low term for the fp8 pipe: the residual x - hi itself (|.| <= half an ulp of hi <= 8) and g - mu'', both e4m3
                const __nv_fp8_storage_t l8 = __nv_cvt_float_to_fp8((float)res, __NV_SATFINITE, __NV_E4M3);
                lut_l8[s][c] = (unsigned char)l8;
                lut_h8[s][c] = (unsigned char)__nv_cvt_float_to_fp8((float)h8v[c], __NV_SATFINITE, __NV_E4M3);
                // rank-one part v_k: delta * B (+ d * lo8 when the low term uses g - mu'' instead of g - mu')
                double uu = delta * x * inv_scale;
                if (p.fp8lo) uu += d8 * (double)__half2float(__half(__nv_cvt_fp8_to_halfraw(l8, __NV_E4M3))) * inv_scale;
                lut_u[s][c] = (float)uu;
            }
        }
    }
    __syncthreads();
    // Each lane owns SNPs 2*lane, 2*lane+1 of the tile, each warp 32 consecutive rows; a warp writes 128 contiguous bytes per row and
    // fp16 plane (64 per fp8 plane).  The loop is instruction-bound (four small stores per row), so the table look-ups are byte
    // permutes -- the four fp16 values of a SNP live in two registers and PRMT picks the pair of bytes of code c (selector
    // 0x10 + 0x22 c); the four fp8 values live in one, and ONE PRMT picks both SNPs' bytes -- and four rows share one code byte.
    const int sa = 2 * lane, sb = 2 * lane + 1;
    auto word_of = [](const __half* h) { return *reinterpret_cast<const uint32_t*>(h); };
    const uint32_t hiA01 = word_of(&lut_hi[sa][0]), hiA23 = word_of(&lut_hi[sa][2]), hiB01 = word_of(&lut_hi[sb][0]), hiB23 = word_of(&lut_hi[sb][2]);
    const uint32_t loA01 = word_of(&lut_lo[sa][0]), loA23 = word_of(&lut_lo[sa][2]), loB01 = word_of(&lut_lo[sb][0]), loB23 = word_of(&lut_lo[sb][2]);
    const uint32_t p2A01 = word_of(&lut_p2[sa][0]), p2A23 = word_of(&lut_p2[sa][2]), p2B01 = word_of(&lut_p2[sb][0]), p2B23 = word_of(&lut_p2[sb][2]);
    // the four fp8 table values of a SNP packed in one word: byte `code`
    const uint32_t l8a = *reinterpret_cast<const uint32_t*>(lut_l8[sa]), l8b = *reinterpret_cast<const uint32_t*>(lut_l8[sb]);
    const uint32_t h8a = *reinterpret_cast<const uint32_t*>(lut_h8[sa]), h8b = *reinterpret_cast<const uint32_t*>(lut_h8[sb]);
    const unsigned char* c0 = codes[sa];
    const unsigned char* c1 = codes[sb];
    auto pick2 = [](uint32_t a01, uint32_t a23, uint32_t b01, uint32_t b23, uint32_t sel_a, uint32_t sel_b) {
        return __byte_perm(__byte_perm(a01, a23, sel_a), __byte_perm(b01, b23, sel_b), 0x5410);     // {SNP a, SNP b} as fp16 x 2
    };
    unsigned char* const l8 = reinterpret_cast<unsigned char*>(p.p2);
    unsigned char* const h8 = l8 + (p.p2_rows ? p.p2_rows : p.n_pad) * p.k_pad;
    for (int g = 0; g < 8; ++g) {
        const int rbase = warp * 32 + g * 4;
        const uint32_t byte_a = c0[rbase >> 2], byte_b = c1[rbase >> 2];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int r = rbase + t;
            const uint32_t ca = (byte_a >> (2 * t)) & 3u, cb = (byte_b >> (2 * t)) & 3u;
            const uint32_t sel_a = 0x10u + 0x22u * ca, sel_b = 0x10u + 0x22u * cb;
            const long long off = (i0 + r) * p.k_pad + b0 + 2 * lane;
            *reinterpret_cast<uint32_t*>(p.hi + off) = pick2(hiA01, hiA23, hiB01, hiB23, sel_a, sel_b);
            *reinterpret_cast<uint32_t*>(p.lo + off) = pick2(loA01, loA23, loB01, loB23, sel_a, sel_b);
            if (fast) {
                // plane 2 may be shared with a second operand stacked above / below this one (train x test): its own row offset
                const long long offp = (p.p2_row0 + i0 + r) * p.k_pad + b0 + 2 * lane;
                if (p.fp8lo) {
                    const uint32_t sel8 = ca | ((cb + 4u) << 4);                                    // byte ca of the first word, byte cb of the second
                    *reinterpret_cast<uint16_t*>(l8 + offp) = (uint16_t)__byte_perm(l8a, l8b, sel8);
                    *reinterpret_cast<uint16_t*>(h8 + offp) = (uint16_t)__byte_perm(h8a, h8b, sel8);
                } else {
                    *reinterpret_cast<uint32_t*>(p.p2 + offp) = pick2(p2A01, p2A23, p2B01, p2B23, sel_a, sel_b);
                }
            }
        }
    }
    // rank-one vector: one thread per row of the tile sums its 64 SNPs' table values (no cross-lane reduction; the table reads are
    // broadcasts -- a warp reads one SNP's four entries -- and the code bytes of 32 rows are 8 adjacent bytes)
    if (fast && p.u) {
        const int r = threadIdx.x;                                      // blockDim.x == PT_I
        float acc = 0.0f;
#pragma unroll 8
        for (int sn = 0; sn < PT_S; ++sn) acc += lut_u[sn][((uint32_t)codes[sn][r >> 2] >> (2 * (r & 3))) & 3u];
        if (acc != 0.0f && i0 + r < n_out) atomicAdd(p.u + i0 + r, (double)acc);
    }
}

// 4 resident CTAs per SM at 64 registers (32 bytes of spilled table registers) or 3 at 80: cfg3 chunk 0.31 vs 0.325 ms
// (PSTB_PLANES_MINB=3: A/B runs)
static void launch_planes(const PlaneParams& pp, long long ptiles, cudaStream_t st) {
    static const int minb = [] { const char* e = getenv("PSTB_PLANES_MINB"); const int v = e ? atoi(e) : 0; return v == 3 ? 3 : 4; }();
    if (minb == 4) k_planes<4><<<(unsigned)ptiles, 256, 0, st>>>(pp);
    else k_planes<3><<<(unsigned)ptiles, 256, 0, st>>>(pp);
}

// ---- tcgen05 / TMA primitives ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16, fp16 A/B (format 0), fp32 accumulate, both K-major, M=128, N=256
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

#define PSTB_TMEM_LD32(taddr, v)                                                                                             \
    asm volatile(                                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                            \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),        \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), \
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),             \
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                        \
        : "r"(taddr)                                                                                                         \
        : "memory")

struct SyrkParams {
    const int2* tiles;      // (I, J): rows [128 I, +128), columns [256 J, +256)
    int ntiles, num_kb;
    float* K;
    long long n, ldk;
    int accumulate;
    const Scalars* sc;      // NULL -> out_scale is used as is
    float out_scale;
    int run_kb_fast;        // k-blocks per TMEM run on the 2-term path (tiles on the diagonal)
    int run_kb_off;         // ... for tiles off the diagonal: their sums have no sign, so the truncating accumulator has no bias to limit
    int compact;            // K is tile storage [ntiles][256][256] (K-tile sharding: this rank's tiles only), 2-CTA kernel only
    // rectangular (train x test) product on stacked planes, 2-CTA kernel only: plane rows [0, row0) hold the column operand,
    // rows [row0, row0 + n) the row operand; tiles are (I, J) with 256 I >= row0 > 256 J; output row = plane row - row0,
    // n_cols columns.  Symmetric product: row0 = 0, n_cols = n.
    long long row0, n_cols;
    int fp8lo;              // 2-term path with the low term on the fp8 pipe: map_p2 / map_h8 are byte (e4m3) planes
    int* counter;           // dynamic tile feed: zeroed device counter of this launch (NULL: static round-robin tile lists)
    int tma_out;            // K leaves through 32 x 32 staging tiles and TMA bulk tensor stores / reduce-adds (map_out)
    int red_add;            // accumulate with red.global.add.v4.f32 (no read round trip in the epilogue) instead of load + add + store
    int dbg;                // timing experiments only (PSTB_SYRK_DBG): bit 1 = no K write at all, bit 2 = no operand loads
};

__global__ void __launch_bounds__(SYRK_THREADS, 1)
k_syrk(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const SyrkParams p) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_hi);
        prefetch_tmap(&map_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 8); }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int num_runs = (p.num_kb + RUN_KB - 1) / RUN_KB;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                const int2 tile = p.tiles[t];
                const int row_a = tile.x * BM, row_b = tile.y * BN;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1u);
                    const uint32_t sb = tiles_base + stage * STAGE_BYTES;
                    const uint32_t full = smem_u32(&bar_full[stage]);
                    mbar_expect_tx(&bar_full[stage], STAGE_BYTES);
                    const int kc = kb * BK;
                    tma_load_2d(sb, &map_hi, kc, row_a, full);
                    tma_load_2d(sb + A_BYTES, &map_lo, kc, row_a, full);
                    tma_load_2d(sb + 2 * A_BYTES, &map_hi, kc, row_b, full);
                    tma_load_2d(sb + 2 * A_BYTES + A_BYTES, &map_hi, kc, row_b + 128, full);
                    tma_load_2d(sb + 2 * A_BYTES + B_BYTES, &map_lo, kc, row_b, full);
                    tma_load_2d(sb + 2 * A_BYTES + B_BYTES + A_BYTES, &map_lo, kc, row_b + 128, full);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, run = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                for (int r = 0; r < num_runs; ++r, ++run) {
                    const uint32_t acc = run & 1u;
                    mbar_wait(&bar_tempty[acc], ((run >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    const int kb_end = min(p.num_kb, (r + 1) * RUN_KB);
                    for (int kb = r * RUN_KB; kb < kb_end; ++kb) {
                        mbar_wait(&bar_full[stage], phase);
                        tc_fence_after();
                        const uint32_t sb = tiles_base + stage * STAGE_BYTES;
                        const uint64_t a_hi = make_smem_desc(sb), a_lo = make_smem_desc(sb + A_BYTES);
                        const uint64_t b_hi = make_smem_desc(sb + 2 * A_BYTES), b_lo = make_smem_desc(sb + 2 * A_BYTES + B_BYTES);
                        const uint32_t first = (kb == r * RUN_KB) ? 0u : 1u;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);   // 16 fp16 = 32 bytes along K inside the swizzle atom
                            // small cross terms first: they are added while the accumulator is still small
                            umma_f16(d_tmem, a_hi + adv, b_lo + adv, kIdesc, (k == 0) ? first : 1u);
                            umma_f16(d_tmem, a_lo + adv, b_hi + adv, kIdesc, 1u);
                            umma_f16(d_tmem, a_hi + adv, b_hi + adv, kIdesc, 1u);
                        }
                        tc_commit(smem_u32(&bar_empty[stage]));
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit(smem_u32(&bar_tfull[acc]));
                }
            }
        }
    } else {
        // ===== epilogue: drain TMEM runs into registers, write K once per tile =====
        const int quad = warp & 3;                                  // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                           // which 128 accumulator columns
        float scale = p.out_scale;
        if (p.sc) scale *= exp2f(-2.0f * (float)scale_exponent(p.sc->absmax_bits));
        const bool vec = (p.ldk % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0);
        uint32_t run = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int2 tile = p.tiles[t];
            float sum[128];
#pragma unroll
            for (int q = 0; q < 128; ++q) sum[q] = 0.0f;
            for (int r = 0; r < num_runs; ++r, ++run) {
                const uint32_t acc = run & 1u;
                mbar_wait(&bar_tfull[acc], (run >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + acc * BN + half * 128 + c * 32 + ((uint32_t)(quad * 32) << 16);
                    PSTB_TMEM_LD32(taddr, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int q = 0; q < 32; ++q) sum[c * 32 + q] += __uint_as_float(v[q]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            }
            const long long row = (long long)tile.x * BM + quad * 32 + lane;
            const long long col0 = (long long)tile.y * BN + half * 128;
            const long long row_hi = (long long)tile.x * BM + quad * 32 + 31;   // last row of this warp
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long cc = col0 + c * 32;
                if (cc > row_hi || cc >= p.n) continue;              // warp-uniform: nothing at or below the diagonal
                if (row < p.n) {
                    float* dst = p.K + row * p.ldk + cc;
                    if (vec && cc + 32 <= p.n) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 o = make_float4(sum[c * 32 + 4 * q] * scale, sum[c * 32 + 4 * q + 1] * scale,
                                                   sum[c * 32 + 4 * q + 2] * scale, sum[c * 32 + 4 * q + 3] * scale);
                            float4* d4 = reinterpret_cast<float4*>(dst) + q;
                            if (p.accumulate) { float4 old = *d4; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                            *d4 = o;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            if (cc + q < p.n) {
                                float o = sum[c * 32 + q] * scale;
                                if (p.accumulate) o += dst[q];
                                dst[q] = o;
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- K3, 2-CTA version: one 256x256 tile per CTA pair (tcgen05 cta_group::2) -----------------------------------
// Each CTA of the pair owns 128 rows of the tile (its 128 TMEM lanes x 256 accumulator columns) and stages its own
// 128 A rows plus HALF of the 256 B rows (hi + lo: 64 KiB per stage, 3 stages); the tensor cores of both SMs read the
// B halves from each other's shared memory.  Per k-block the pair loads 128 KiB for 25 MFLOP (196 flop/B, vs 128 for the
// 1-CTA 128x256 tile) and every SM reads a third less shared memory per MMA.  Protocol:
//   full[stage]   lives in the leader (rank 0): leader arms 128 KiB expect-tx, both CTAs' TMA loads complete on it;
//   empty[stage]  in both CTAs, released by a multicast tcgen05.commit from the leader's MMA thread;
//   tfull[acc]    in both CTAs (multicast commit); tempty[acc] in the leader, 16 arrivals (8 epilogue warps x 2 CTAs).
namespace v2 {
constexpr int TM = 256, TN = 256;                          // tile of the pair
constexpr int STAGES2 = 3;
constexpr int T_BYTES = 128 * BK * 2;                      // 16 KiB: one 128-row fp16 plane tile
constexpr int STAGE2_BYTES = 4 * T_BYTES;                  // A hi, A lo, B-half hi, B-half lo
constexpr int OUT_STAGE_BYTES = 32 * 32 * 4;                // one 32 x 32 fp32 block per epilogue warp, staged for the TMA store of K
constexpr int SYRK2_SMEM = STAGES2 * STAGE2_BYTES + 8 * OUT_STAGE_BYTES + 1024;
constexpr uint32_t kIdesc2 = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr int GROUP2 = 8;                                  // 8 x 8 tiles = 2048 x 2048 super-blocks

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_mc2(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major byte tile, 64-byte swizzle: rows of 64 bytes, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_addr, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_addr),
        "r"(cta)
        : "memory");
}

constexpr int SCHED_SLOTS = 4;
__device__ __forceinline__ void st_shared_remote(uint32_t local_addr, uint32_t cta, uint32_t value) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.u32 [ra], %2;\n\t}" ::"r"(local_addr),
        "r"(cta), "r"(value)
        : "memory");
}
// wait with cluster-scope acquire: the data the barrier guards may have been written by the peer CTA (distributed shared memory)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    for (uint32_t spin = 1;; ++spin) {
        if (mbar_try_wait_cluster(bar, parity)) return;
        if ((spin & 1023u) == 0 && global_ns() - t0 > 4000000000ull) __trap();
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SYRK_THREADS, 1)
k_syrk2(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_p2,
        const __grid_constant__ CUtensorMap map_h8, const __grid_constant__ CUtensorMap map_out, const SyrkParams p) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[STAGES2 + 1], bar_empty[STAGES2 + 1], bar_tfull[2], bar_tempty[2];
    __shared__ __align__(8) uint64_t bar_sfull[SCHED_SLOTS], bar_sempty[SCHED_SLOTS];
    __shared__ int sched_tile[SCHED_SLOTS];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    // ---- tile feed ---------------------------------------------------------------------------------------------------------------
    // Static (p.counter == NULL): CTA pair c takes tiles c, c + pairs, ... .  Dynamic: the leader's producer thread draws the next tile
    // index from a global counter and publishes it to every role of both CTAs through a small ring (value + full / empty mbarriers;
    // the peer's copy is written through distributed shared memory).  A pair that gets its SMs late -- a collective or another kernel
    // holds them when the launch starts -- then simply draws fewer tiles; with the static lists it would still owe its whole share and
    // the launch would end that much later (measured on 2 GPUs: an all-reduce next to the static kernel cost more than it hid).
    const bool dynamic = p.counter != nullptr;
    auto feed_next = [&](uint32_t& it) -> int {                        // every role but the scheduler; all lanes of a warp may call it
        if (!dynamic) {
            const long long t = (long long)cluster_id + (long long)it * nclusters;
            ++it;
            return t < p.ntiles ? (int)t : -1;
        }
        const uint32_t slot = it % SCHED_SLOTS, ph = (it / SCHED_SLOTS) & 1u;
        mbar_wait_cluster(&bar_sfull[slot], ph);
        const int t = *reinterpret_cast<volatile int*>(&sched_tile[slot]);
        __syncwarp(__activemask());                                    // whole epilogue warps or a single elected lane
        if (lane == 0) {                                               // the slot may be reused once every role has read it
            if (leader) mbar_arrive(&bar_sempty[slot]); else mbar_arrive_remote(smem_u32(&bar_sempty[slot]), 0u);
        }
        ++it;
        return t < p.ntiles ? t : -1;
    };
    auto sched_next = [&](uint32_t& it) -> int {                       // leader CTA, warp 0, lane 0 only
        if (!dynamic) return feed_next(it);
        const uint32_t slot = it % SCHED_SLOTS, ph = (it / SCHED_SLOTS) & 1u;
        mbar_wait(&bar_sempty[slot], ph ^ 1u);
        const int t = atomicAdd(p.counter, 1);
        sched_tile[slot] = t;
        st_shared_remote(smem_u32(&sched_tile[slot]), 1u, (uint32_t)t);
        mbar_arrive(&bar_sfull[slot]);
        mbar_arrive_remote(smem_u32(&bar_sfull[slot]), 1u);
        ++it;
        return t < p.ntiles ? t : -1;
    };
    const uint32_t tiles_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_hi);
        prefetch_tmap(&map_lo);
        prefetch_tmap(&map_p2);
        prefetch_tmap(&map_h8);
        if (p.tma_out) prefetch_tmap(&map_out);
    }
    // exact-dosage chunk (no missing data): A = h (plane 0), B = (w h)_hi, (w h)_lo (planes 1, 2): two MMAs per k-step
    const bool fast = p.sc != nullptr && p.sc->need_3term == 0u;
    // fp8lo: the low term h * (w h)_lo runs on the fp8 pipe (e4m3 x e4m3, K = 32 per instruction: half the tensor cycles of the
    // fp16 low term).  It only has to carry 4-5 bits: (w h)_lo is <= 2^-11 of (w h)_hi.  map_p2 / map_h8 are byte planes then.
    const bool fp8lo = fast && p.fp8lo != 0;
    // the 2-term loop spends a third less time per k-block, so it gets a fourth (smaller) stage from the same 192 KiB ring
    const uint32_t nstages = fast ? STAGES2 + 1 : STAGES2, stage_bytes = fast ? 3u * T_BYTES : STAGE2_BYTES;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES2 + 1; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 16); }
        // tile feed: one arrival publishes a slot; 18 readers free it (leader: MMA thread + 8 epilogue warps; peer: producer + 8 epilogue warps)
        for (int q = 0; q < SCHED_SLOTS; ++q) { mbar_init(&bar_sfull[q], 1); mbar_init(&bar_sempty[q], 18); }
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // k-blocks per TMEM run.  The tensor core truncates (rounds toward zero) on every accumulate: every partial sum loses magnitude,
    // ~2.3e-7 of it per k-block of the run (fitted on cfg3: 6 / 63 k-blocks per run -> 5.3e-6 / 1.5e-5 worst off-diagonal tile, fp8 low
    // term).  On the diagonal of K -- long positive sums that carry most of the Frobenius norm -- that is a bias, so diagonal tiles drain
    // their accumulator every 6 k-blocks.  Every drain idles the tensor pipe for ~500 cycles (6 / 12 / 24 / 63 k-blocks per run =
    // 1485 / 1532 / 1561 / 1624 TFLOP/s on the cfg3 shape), so off-diagonal tiles (99 % of them) run 12 k-blocks: +3 % for ~2.8e-6 of
    // shrink on entries whose error budget is dominated by the fp8 low term (5e-6).  One run per chunk would be +9 % but 1.5e-5.
    const int run_kb_diag = fast ? p.run_kb_fast : RUN_KB;
    const int run_kb_offd = (fast && p.run_kb_off > run_kb_diag) ? p.run_kb_off : run_kb_diag;
    auto run_kb_of = [&](const int2& tile) { return (tile.x == tile.y) ? run_kb_diag : run_kb_offd; };

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, feed = 0;
            for (int t = leader ? sched_next(feed) : feed_next(feed); t >= 0; t = leader ? sched_next(feed) : feed_next(feed)) {
                const int2 tile = p.tiles[t];
                const int row_a = tile.x * TM + (int)rank * 128, row_b = tile.y * TN + (int)rank * 128;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1u);
                    const uint32_t sb = tiles_base + stage * stage_bytes;
                    const uint32_t full = smem_u32(&bar_full[stage]) & 0xFEFFFFFFu;   // the leader CTA's barrier
                    const int kc = kb * BK;
                    if (p.dbg & 4) {                                              // timing experiment: MMAs on whatever is in smem
                        if (leader) mbar_arrive(&bar_full[stage]);
                        if (++stage == nstages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    if (fast) {
                        if (leader) mbar_expect_tx(&bar_full[stage], 2u * 3u * T_BYTES);
                        tma_load_2d_2sm(sb, &map_hi, kc, row_a, full);
                        tma_load_2d_2sm(sb + T_BYTES, &map_lo, kc, row_b, full);
                        if (fp8lo) {
                            tma_load_2d_2sm(sb + 2 * T_BYTES, &map_h8, kc, row_a, full);                 // h, e4m3: 8 KiB
                            tma_load_2d_2sm(sb + 2 * T_BYTES + T_BYTES / 2, &map_p2, kc, row_b, full);   // (w h)_lo, e4m3: 8 KiB
                        } else {
                            tma_load_2d_2sm(sb + 2 * T_BYTES, &map_p2, kc, row_b, full);
                        }
                    } else {
                        if (leader) mbar_expect_tx(&bar_full[stage], 2u * STAGE2_BYTES);
                        tma_load_2d_2sm(sb, &map_hi, kc, row_a, full);
                        tma_load_2d_2sm(sb + T_BYTES, &map_lo, kc, row_a, full);
                        tma_load_2d_2sm(sb + 2 * T_BYTES, &map_hi, kc, row_b, full);
                        tma_load_2d_2sm(sb + 3 * T_BYTES, &map_lo, kc, row_b, full);
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread of the leader CTA) =====
        if (leader && lane == 0) {
            uint32_t stage = 0, phase = 0, run = 0, feed = 0;
            for (int t = feed_next(feed); t >= 0; t = feed_next(feed)) {
                const int run_kb = run_kb_of(p.tiles[t]);
                const int num_runs = (p.num_kb + run_kb - 1) / run_kb;
                for (int r = 0; r < num_runs; ++r, ++run) {
                    const uint32_t acc = run & 1u;
                    mbar_wait(&bar_tempty[acc], ((run >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TN;
                    const int kb_end = min(p.num_kb, (r + 1) * run_kb);
                    for (int kb = r * run_kb; kb < kb_end; ++kb) {
                        mbar_wait(&bar_full[stage], phase);
                        tc_fence_after();
                        const uint32_t sb = tiles_base + stage * stage_bytes;
                        // 3-term: [x_hi | x_lo | B x_hi | B x_lo];  2-term: [h | B (w h)_hi | B (w h)_lo]
                        const uint64_t a_hi = make_smem_desc(sb), a_lo = make_smem_desc(sb + T_BYTES);
                        const uint64_t b_hi = make_smem_desc(sb + (fast ? 1 : 2) * T_BYTES), b_lo = make_smem_desc(sb + (fast ? 2 : 3) * T_BYTES);
                        const uint32_t first = (kb == r * run_kb) ? 0u : 1u;
                        if (fp8lo) {
                            // slots: [h fp16 | B (w h)_hi fp16 | h e4m3 (8 KiB) | B (w h)_lo e4m3 (8 KiB)]
                            const uint64_t a8 = make_smem_desc_sw64(sb + 2 * T_BYTES), b8 = make_smem_desc_sw64(sb + 2 * T_BYTES + T_BYTES / 2);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                umma_f16_2sm(d_tmem, a_hi + adv, b_hi + adv, kIdesc2, (k == 0) ? first : 1u);
                            }
#pragma unroll
                            for (int k = 0; k < BK / 32; ++k) {
                                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                                umma_f8_2sm(d_tmem, a8 + adv, b8 + adv, kIdesc2, 1u);
                            }
                            tc_commit_mc2(smem_u32(&bar_empty[stage]));
                            if (++stage == nstages) { stage = 0; phase ^= 1u; }
                            continue;
                        }
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            if (fast) {
                                // slots: a_hi = h, b_hi = (w h)_hi, b_lo = (w h)_lo
                                umma_f16_2sm(d_tmem, a_hi + adv, b_lo + adv, kIdesc2, (k == 0) ? first : 1u);
                                umma_f16_2sm(d_tmem, a_hi + adv, b_hi + adv, kIdesc2, 1u);
                            } else {
                                umma_f16_2sm(d_tmem, a_hi + adv, b_lo + adv, kIdesc2, (k == 0) ? first : 1u);
                                umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, kIdesc2, 1u);
                                umma_f16_2sm(d_tmem, a_hi + adv, b_hi + adv, kIdesc2, 1u);
                            }
                        }
                        tc_commit_mc2(smem_u32(&bar_empty[stage]));
                        if (++stage == nstages) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit_mc2(smem_u32(&bar_tfull[acc]));
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs): drain TMEM runs into registers, write K once per tile =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        float scale = p.out_scale;
        if (fast) scale *= exp2f(-(float)scale_exponent(p.sc->absmax_w_bits));            // only the right operand is scaled
        else if (p.sc) scale *= exp2f(-2.0f * (float)scale_exponent(p.sc->absmax_bits));
        const bool vec = (p.ldk % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0);
        uint32_t run = 0, feed = 0;
        for (int t = feed_next(feed); t >= 0; t = feed_next(feed)) {
            const int2 tile = p.tiles[t];
            const int run_kb = run_kb_of(tile);
            const int num_runs = (p.num_kb + run_kb - 1) / run_kb;
            float sum[128];
#pragma unroll
            for (int q = 0; q < 128; ++q) sum[q] = 0.0f;
            for (int r = 0; r < num_runs; ++r, ++run) {
                const uint32_t acc = run & 1u;
                mbar_wait(&bar_tfull[acc], (run >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + acc * TN + half * 128 + c * 32 + ((uint32_t)(quad * 32) << 16);
                    PSTB_TMEM_LD32(taddr, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int q = 0; q < 32; ++q) sum[c * 32 + q] += __uint_as_float(v[q]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(smem_u32(&bar_tempty[acc]), 0u);
            }
            const long long row = (long long)tile.x * TM + rank * 128 + quad * 32 + lane - p.row0;
            const long long col0 = (long long)tile.y * TN + half * 128;
            const long long row_hi = (long long)tile.x * TM + rank * 128 + quad * 32 + 31;       // plane row: never above a stacked column
            // compact: tile t of this rank's list is stored whole (256 x 256, ld 256), diagonal tiles included in full
            float* kbase = p.compact ? p.K + (long long)t * (TM * TN) + (long long)(rank * 128 + quad * 32 + lane) * TN + half * 128
                                     : p.K + row * p.ldk + col0;
            const bool vec_t = p.compact || vec;
            if (p.tma_out) {
                // K (+)= block through shared memory and the TMA: the warp's 32 x 32 block (lane = row) is written to a 128-byte-swizzled
                // staging tile (conflict-free 128-bit stores) and leaves as ONE bulk tensor store -- or reduce-add, performed at L2 --
                // instead of 256 scattered 16-byte accesses per warp (REDG costs ~1.3 cycles per lane: ~10 k cycles per tile, during
                // which the MMA thread ran out of accumulators).  Rows / columns beyond the matrix are clipped by the tensor map.
                const uint32_t stage_out = tiles_base + STAGES2 * STAGE2_BYTES + (uint32_t)(warp - 2) * OUT_STAGE_BYTES;
                const int out_row = p.compact ? t * TM + (int)rank * 128 + quad * 32
                                              : (int)((long long)tile.x * TM + rank * 128 + quad * 32 - p.row0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const long long cc = col0 + c * 32;
                    if ((!p.compact && cc > row_hi) || cc >= p.n_cols || (p.dbg & 2)) continue;       // warp-uniform
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous block has left the staging tile
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t dst = stage_out + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(sum[c * 32 + 4 * q] * scale),
                                     "f"(sum[c * 32 + 4 * q + 1] * scale), "f"(sum[c * 32 + 4 * q + 2] * scale), "f"(sum[c * 32 + 4 * q + 3] * scale)
                                     : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int oc = p.compact ? half * 128 + c * 32 : (int)cc;
                        if (p.accumulate)
                            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_out),
                                         "r"(stage_out), "r"(oc), "r"(out_row)
                                         : "memory");
                        else
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_out),
                                         "r"(stage_out), "r"(oc), "r"(out_row)
                                         : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                continue;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long cc = col0 + c * 32;
                if ((!p.compact && cc > row_hi) || cc >= p.n_cols) continue;
                if (row < p.n) {
                    float* dst = kbase + c * 32;
                    if (p.dbg & 2) continue;                                       // timing experiment: no K traffic
                    if (vec_t && cc + 32 <= p.n_cols) {
                        if (p.accumulate && p.red_add) {
                            // K += sum as a fire-and-forget vector reduction performed at L2: the epilogue warps do not wait for
                            // a DRAM round trip per 32 columns, so the MMA thread gets its accumulator back sooner
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(sum[c * 32 + 4 * q] * scale),
                                             "f"(sum[c * 32 + 4 * q + 1] * scale), "f"(sum[c * 32 + 4 * q + 2] * scale),
                                             "f"(sum[c * 32 + 4 * q + 3] * scale)
                                             : "memory");
                            continue;
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 o = make_float4(sum[c * 32 + 4 * q] * scale, sum[c * 32 + 4 * q + 1] * scale,
                                                   sum[c * 32 + 4 * q + 2] * scale, sum[c * 32 + 4 * q + 3] * scale);
                            float4* d4 = reinterpret_cast<float4*>(dst) + q;
                            if (p.accumulate) { float4 old = *d4; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                            *d4 = o;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            if (cc + q < p.n_cols) {
                                float o = sum[c * 32 + q] * scale;
                                if (p.accumulate) o += dst[q];
                                dst[q] = o;
                            }
                        }
                    }
                }
            }
        }
    }
    if (p.tma_out && warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

void build_tiles2(long long n, std::vector<int2>& out, int group = GROUP2) {
    out.clear();
    const int t = (int)((n + TM - 1) / TM);
    for (int gi = 0; gi < t; gi += group)
        for (int gj = 0; gj <= gi; gj += group)
            for (int I = gi; I < gi + group && I < t; ++I)
                for (int J = gj; J < gj + group && J <= I; ++J) out.push_back(make_int2(I, J));
}
}  // namespace v2

// exact-dosage path: K_ik += v_k on the lower triangle (once per call, before the mirror)
__global__ void __launch_bounds__(256) k_apply_rank1(float* K, long long n, long long ldk, const double* u) {
    const long long i = blockIdx.x;
    for (long long k = (long long)blockIdx.y * blockDim.x + threadIdx.x; k <= i; k += (long long)gridDim.y * blockDim.x)
        K[i * ldk + k] = (float)((double)K[i * ldk + k] + u[k]);
}
// the same on compact tile storage [ntiles][256][256]
__global__ void __launch_bounds__(256) k_apply_rank1_tiles(float* tiles, const int2* coords, long long n, const double* u) {
    const int2 t = coords[blockIdx.x];
    float* base = tiles + (long long)blockIdx.x * 65536;
    for (int e = threadIdx.x; e < 65536; e += blockDim.x) {
        const long long i = (long long)t.x * 256 + (e >> 8), k = (long long)t.y * 256 + (e & 255);
        if (i < n && k < n) base[e] = (float)((double)base[e] + u[k]);
    }
}
// the same for a rectangular (train x test) product: out[i, k] += v_k
__global__ void __launch_bounds__(256) k_apply_colvec(float* out, long long n_rows, long long n_cols, long long ldo, const double* u) {
    const long long i = blockIdx.x;
    for (long long k = (long long)blockIdx.y * blockDim.x + threadIdx.x; k < n_cols; k += (long long)gridDim.y * blockDim.x)
        out[i * ldo + k] = (float)((double)out[i * ldo + k] + u[k]);
}

__global__ void __launch_bounds__(256) k_mirror(float* K, long long n, long long ldk) {
    __shared__ float tile[32][33];
    const long long bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long i = bi * 32 + r, j = bj * 32 + tx;
        tile[r][tx] = (i < n && j < n) ? K[i * ldk + j] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long j = bj * 32 + r, i = bi * 32 + tx;     // writes K[j][i] = K[i][j] for j < i
        if (i < n && j < n && j < i) K[j * ldk + i] = tile[tx][r];
    }
}

// compact tile storage [ntiles][256][256] (lower-triangular tiles of `coords`) -> full K, both triangles.  One CTA per 32 x 32
// sub-block: a coalesced copy into the lower triangle and, through shared memory, its transpose into the upper one.  Of a
// diagonal tile only the lower half is used, so K comes out exactly symmetric.
// u (optional): the rank-one part of the exact-dosage path, K_ik += u_k, added on the way (deferred from the SYRK calls so that the
// multi-GPU path all-reduces it as one small vector instead of sweeping the triangle once more per call)
__global__ void __launch_bounds__(256) k_untile(const float* tiles, const int2* coords, long long n, float* K, long long ldk, const double* u) {
    __shared__ float sub[32][33];
    const int2 t = coords[blockIdx.x];
    const int sr = blockIdx.y >> 3, sc = blockIdx.y & 7;
    if (t.x == t.y && sc > sr) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* src = tiles + (long long)blockIdx.x * 65536 + (long long)(sr * 32) * 256 + sc * 32;
    const long long i0 = (long long)t.x * 256 + sr * 32, k0 = (long long)t.y * 256 + sc * 32;
    const double uk = (u && k0 + tx < n) ? u[k0 + tx] : 0.0;
    for (int r = ty; r < 32; r += 8) {
        const long long i = i0 + r, k = k0 + tx;
        const float v = u ? (float)((double)src[r * 256 + tx] + uk) : src[r * 256 + tx];
        sub[r][tx] = v;
        if (i < n && k < n && k <= i) K[i * ldk + k] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long k = k0 + r, i = i0 + tx;                 // writes K[k][i] = K[i][k] for k < i
        if (i < n && k < n && k < i) K[k * ldk + i] = sub[tx][r];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_convert(const float* K, long long total, T* out, double scale) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        out[e] = (T)((double)K[e] * scale);
}

// ---- float matrices (SnpData._read_kernel: val.dot(val.T), snpdata.py:203-206) -> operand planes ----------------
template <typename T>
__global__ void __launch_bounds__(256) k_absmax_float(const T* val, long long total, Scalars* sc) {
    float m = 0.0f;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        double v = fabs((double)val[e]);
        if (v == v && v < 1e300) m = fmaxf(m, __double2float_ru(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(&sc->absmax_bits, __float_as_uint(m));
}

// val[i * si + j * sj], columns [c0, c0 + ns) -> hi/lo planes [n_pad][k_pad] (zero padded), 32 x 32 tiles through smem
template <typename T>
__global__ void __launch_bounds__(256) k_split_planes(const T* val, long long si, long long sj, long long n, long long c0, long long ns,
                                                      const Scalars* sc, __half* hi, __half* lo, long long n_pad, long long k_pad) {
    __shared__ float th[32][33], tl[32][33];
    const long long ti = blockIdx.y, tj = blockIdx.x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double scale = ldexp(1.0, scale_exponent(sc->absmax_bits));
    const bool row_fast = (si == 1);                       // F order: consecutive individuals are adjacent in memory
    for (int r = ty; r < 32; r += 8) {
        const long long i = ti * 32 + (row_fast ? tx : r), j = tj * 32 + (row_fast ? r : tx);
        double x = 0.0;
        if (i < n && j < ns) x = (double)val[i * si + (c0 + j) * sj] * scale;
        if (fabs(x) > 60000.0) x = 0.0;
        const __half h = __float2half_rn((float)x);
        const float hf = __half2float(h), lf = (float)(x - (double)hf);
        if (row_fast) { th[r][tx] = hf; tl[r][tx] = lf; }       // [j_local][i_local]
        else { th[tx][r] = hf; tl[tx][r] = lf; }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long i = ti * 32 + r, j = tj * 32 + tx;      // write with the SNP index fastest
        if (i < n_pad && j < k_pad) {
            hi[i * k_pad + j] = __float2half_rn(th[tx][r]);
            lo[i * k_pad + j] = __float2half_rn(tl[tx][r]);
        }
    }
}

// ---- host helpers --------------------------------------------------------------------------------------------
// Tuning knobs, read from the environment on every launch (kernel experiments; the defaults are the shipped configuration).
struct Knobs {
    int run_kb_fast;   // PSTB_RUN_KB_FAST: k-blocks per TMEM run on the 2-term path, diagonal tiles (default 6)
    int run_kb_off;    // PSTB_RUN_KB_OFF: ... off-diagonal tiles (default 12)
    int tma_out;       // PSTB_SYRK_TMA_OUT: 1 = K leaves through staging tiles + TMA bulk tensor store / reduce-add (default), 0 = per-thread path
    int red_add;       // PSTB_SYRK_RED: per-thread path: 1 = red.global.add.v4.f32 (default), 0 = load + add + store
    int dbg;           // PSTB_SYRK_DBG: timing experiments only (results are wrong): 2 = no K write, 4 = no operand loads
    int clusters;      // PSTB_SYRK_CLUSTERS: CTA pairs to launch (default: SM count / 2)
    int dynamic;       // PSTB_SYRK_DYN: 1 = dynamic tile feed (default), 0 = static round-robin tile lists
    int group;         // PSTB_SYRK_GROUP: tiles per side of a rasterisation super-block (default 8 = 2048 x 2048)
};
Knobs knobs() {
    auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return (e && *e) ? atoi(e) : dflt; };
    Knobs k;
    k.run_kb_fast = geti("PSTB_RUN_KB_FAST", 6);
    if (k.run_kb_fast < 1 || k.run_kb_fast > 64) k.run_kb_fast = 6;
    k.run_kb_off = geti("PSTB_RUN_KB_OFF", 12);
    if (k.run_kb_off < 1) k.run_kb_off = 12;
    k.tma_out = geti("PSTB_SYRK_TMA_OUT", 1) != 0 ? 1 : 0;
    k.red_add = geti("PSTB_SYRK_RED", 1) != 0 ? 1 : 0;
    k.dbg = geti("PSTB_SYRK_DBG", 0);
    k.clusters = geti("PSTB_SYRK_CLUSTERS", 0);
    k.dynamic = geti("PSTB_SYRK_DYN", 1) != 0 ? 1 : 0;
    k.group = geti("PSTB_SYRK_GROUP", 8);
    if (k.group < 1 || k.group > 64) k.group = 8;
    return k;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int make_plane_map(CUtensorMap* map, const void* plane, long long n_pad, long long k_pad) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)k_pad * 2};
    cuuint32_t box[2] = {BK, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(plane), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

// byte (e4m3) plane [n_pad][k_pad]: boxes of 64 bytes x 128 rows, 64-byte swizzle
int make_byte_plane_map(CUtensorMap* map, const void* plane, long long n_pad, long long k_pad) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)k_pad};
    cuuint32_t box[2] = {BK, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(plane), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (byte plane) failed with code %d", (int)r);
    return 0;
}

// float32 output matrix [rows][cols] with leading dimension ld (elements): boxes of 32 x 32, 128-byte swizzle (the epilogue's staging
// tiles).  Returns 0 and sets *ok = 1 when the matrix can be addressed by a tensor map (16-byte aligned base and row pitch).
int make_out_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int* ok) {
    *ok = 0;
    memset(map, 0, sizeof(*map));
    if (rows < 1 || cols < 1 || (ld % 4) != 0 || (reinterpret_cast<uintptr_t>(base) & 15u) != 0 || rows > 0x7fffffffLL || cols > 0x7fffffffLL) return 0;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return 0;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) *ok = 1;
    return 0;
}

// lower-triangular tile list, rasterised in GROUP_I x GROUP_J super-blocks so concurrently running CTAs share operand rows
void build_tiles(long long n, std::vector<int2>& out) {
    out.clear();
    const int ti = (int)((n + BM - 1) / BM), tj = (int)((n + BN - 1) / BN);
    for (int gi = 0; gi < ti; gi += GROUP_I)
        for (int gj = 0; gj < tj; gj += GROUP_J)
            for (int I = gi; I < gi + GROUP_I && I < ti; ++I)
                for (int J = gj; J < gj + GROUP_J && J < tj; ++J)
                    if ((long long)J * BN <= (long long)I * BM + BM - 1) out.push_back(make_int2(I, J));
}

struct TileCache {
    long long n = -1;
    int device = -1;
    int version = 0, rank = 0, world = 1, group = 0;
    int2* d_tiles = nullptr;
    int ntiles = 0;
};

// tiles of the lower triangle owned by `rank` of `world`: every world-th tile of the rasterised list (balanced, and the
// tiles a rank works on concurrently stay inside one super-block)
void owned_tiles(long long n, int version, int rank, int world, std::vector<int2>& out) {
    std::vector<int2> all;
    if (version == 2) v2::build_tiles2(n, all, knobs().group); else build_tiles(n, all);
    if (world <= 1) { out.swap(all); return; }
    out.clear();
    for (size_t t = (size_t)rank; t < all.size(); t += (size_t)world) out.push_back(all[t]);
}

int get_tiles(long long n, int version, int rank, int world, cudaStream_t st, const int2** d_tiles, int* ntiles) {
    static thread_local TileCache c;
    int dev = 0;
    PSTB_CUDA(cudaGetDevice(&dev));
    const int group = knobs().group;
    if (c.n != n || c.device != dev || c.version != version || c.rank != rank || c.world != world || c.group != group) {
        std::vector<int2> tiles;
        owned_tiles(n, version, rank, world, tiles);
        PSTB_CUDA(cudaStreamSynchronize(st));
        if (c.d_tiles) cudaFree(c.d_tiles);
        c.d_tiles = nullptr;
        c.n = -1;
        PSTB_CUDA(cudaMalloc(&c.d_tiles, (tiles.size() + 1) * sizeof(int2)));
        PSTB_CUDA(cudaMemcpy(c.d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice));
        c.ntiles = (int)tiles.size();
        c.n = n;
        c.device = dev;
        c.version = version;
        c.rank = rank;
        c.world = world;
        c.group = group;
    }
    *d_tiles = c.d_tiles;
    *ntiles = c.ntiles;
    return 0;
}

long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

int launch_syrk(const __half* hi, const __half* lo, long long n, long long n_pad, long long k_pad, float* K, long long ldk,
                int accumulate, const Scalars* sc, float out_scale, cudaStream_t st, int rank = 0, int world = 1, int compact = 0,
                const __half* p2 = nullptr, const int2** tiles_out = nullptr, int* ntiles_out = nullptr, int fp8lo = 0,
                long long tile_begin = 0, long long tile_end = -1, int reserve_sms = 0) {
    if (n_pad % ROW_PAD || k_pad % BK || n_pad < n) return fail("planes must be padded to %d rows / %d columns", ROW_PAD, BK);
    if ((reinterpret_cast<uintptr_t>(hi) & 127u) || (reinterpret_cast<uintptr_t>(lo) & 127u)) return fail("planes must be 128-byte aligned");
    CUtensorMap map_hi, map_lo, map_p2, map_h8;
    if (make_plane_map(&map_hi, hi, n_pad, k_pad) || make_plane_map(&map_lo, lo, n_pad, k_pad)) return 1;
    if (fp8lo && p2) {
        // the third plane's bytes hold two e4m3 planes: (w h)_lo, then h
        const unsigned char* l8 = reinterpret_cast<const unsigned char*>(p2);
        if (make_byte_plane_map(&map_p2, l8, n_pad, k_pad) || make_byte_plane_map(&map_h8, l8 + n_pad * k_pad, n_pad, k_pad)) return 1;
    } else {
        fp8lo = 0;
        if (make_plane_map(&map_p2, p2 ? p2 : lo, n_pad, k_pad)) return 1;
        map_h8 = map_p2;
    }
    static const int env_version = (getenv("PSTB_SYRK_V1") && atoi(getenv("PSTB_SYRK_V1")) != 0) ? 1 : 2;   // 1-CTA 128x256 kept for A/B runs
    const int version = compact ? 2 : env_version;
    const int2* d_tiles = nullptr;
    int ntiles = 0;
    if (get_tiles(n, version, rank, world, st, &d_tiles, &ntiles)) return 1;
    if (tile_end >= 0) {
        // a band of the (compact) tile list: tiles [tile_begin, tile_end) of this rank, stored at their usual place
        if (!compact || tile_begin < 0 || tile_end > ntiles || tile_begin > tile_end) return fail("bad tile range [%lld, %lld) of %d", tile_begin, tile_end, ntiles);
        d_tiles += tile_begin;
        K += tile_begin * (long long)(v2::TM * v2::TN);
        ntiles = (int)(tile_end - tile_begin);
    }
    if (tiles_out) *tiles_out = d_tiles;
    if (ntiles_out) *ntiles_out = ntiles;
    SyrkParams p{};
    p.tiles = d_tiles;
    p.ntiles = ntiles;
    p.num_kb = (int)(k_pad / BK);
    p.K = K;
    p.n = n;
    p.n_cols = n;
    p.row0 = 0;
    p.ldk = ldk;
    p.accumulate = accumulate;
    p.sc = sc;
    p.out_scale = out_scale;
    p.compact = compact;
    p.fp8lo = fp8lo;
    const Knobs kn = knobs();
    p.run_kb_fast = kn.run_kb_fast;
    p.run_kb_off = kn.run_kb_off;
    p.red_add = kn.red_add;
    p.dbg = kn.dbg;
    CUtensorMap map_out;
    int out_ok = 0;
    if (compact) make_out_map(&map_out, K, (long long)ntiles * v2::TM, v2::TN, v2::TN, &out_ok);
    else make_out_map(&map_out, K, n, n, ldk, &out_ok);
    p.tma_out = (kn.tma_out && out_ok) ? 1 : 0;
    // per launch, not cached: the attribute belongs to the (function, device) pair and a thread may switch devices
    PSTB_CUDA(cudaFuncSetAttribute(k_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM));
    PSTB_CUDA(cudaFuncSetAttribute(v2::k_syrk2, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SYRK2_SMEM));
    if (ntiles < 1) return 0;
    if (version == 2) {
        int clusters = sm_count_cached() / 2;
        if (kn.clusters > 0 && kn.clusters < clusters) clusters = kn.clusters;
        if (reserve_sms > 0) clusters = clusters - (reserve_sms + 1) / 2 > 8 ? clusters - (reserve_sms + 1) / 2 : 8;   // leave SMs to a concurrent collective
        if (clusters > ntiles) clusters = ntiles;
        if (kn.dynamic && next_counter(st, &p.counter)) return 1;
        v2::k_syrk2<<<2 * clusters, SYRK_THREADS, v2::SYRK2_SMEM, st>>>(map_hi, map_lo, map_p2, map_h8, map_out, p);   // __cluster_dims__(2,1,1)
        PSTB_AFTER_LAUNCH("k_syrk2");
        return 0;
    }
    int grid = sm_count_cached();
    if (grid > ntiles) grid = ntiles;
    k_syrk<<<grid, SYRK_THREADS, SYRK_SMEM, st>>>(map_hi, map_lo, p);
    PSTB_AFTER_LAUNCH("k_syrk");
    return 0;
}

// ---- rectangular product (train x test) on stacked planes ----------------------------------------------------------------
struct CrossTileCache {
    long long rb = -1, cb = -1;
    int device = -1;
    int2* d_tiles = nullptr;
    int ntiles = 0;
};

// tiles (I, J): I over the row operand's 256-row blocks (stacked after the cb column blocks), J over the column blocks;
// rasterised in GROUP2 x GROUP2 super-blocks like the triangular list
int get_cross_tiles(long long row_blocks, long long col_blocks, cudaStream_t st, const int2** d_tiles, int* ntiles) {
    static thread_local CrossTileCache c;
    int dev = 0;
    PSTB_CUDA(cudaGetDevice(&dev));
    if (c.rb != row_blocks || c.cb != col_blocks || c.device != dev) {
        std::vector<int2> tiles;
        for (long long gi = 0; gi < row_blocks; gi += v2::GROUP2)
            for (long long gj = 0; gj < col_blocks; gj += v2::GROUP2)
                for (long long I = gi; I < gi + v2::GROUP2 && I < row_blocks; ++I)
                    for (long long J = gj; J < gj + v2::GROUP2 && J < col_blocks; ++J)
                        tiles.push_back(make_int2((int)(col_blocks + I), (int)J));
        PSTB_CUDA(cudaStreamSynchronize(st));
        if (c.d_tiles) cudaFree(c.d_tiles);
        c.d_tiles = nullptr;
        c.rb = -1;
        PSTB_CUDA(cudaMalloc(&c.d_tiles, (tiles.size() + 1) * sizeof(int2)));
        PSTB_CUDA(cudaMemcpy(c.d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice));
        c.ntiles = (int)tiles.size();
        c.rb = row_blocks;
        c.cb = col_blocks;
        c.device = dev;
    }
    *d_tiles = c.d_tiles;
    *ntiles = c.ntiles;
    return 0;
}

// out[n_rows, n_cols] (+)= X_rows X_cols^T: planes hold the column operand in rows [0, cols_pad) and the row operand in rows
// [cols_pad, cols_pad + rows_pad).  Same paths as the symmetric kernel (sc->need_3term chooses on the device): on the 2-term path
// the row side is read from plane 0 (L, and its e4m3 twin), the column side from planes 1 / 2 (B hi / lo).
int launch_cross(const __half* hi, const __half* lo, const __half* p2, int fp8lo, long long n_rows, long long rows_pad, long long n_cols,
                 long long cols_pad, long long k_pad, float* out, long long ldo, int accumulate, const Scalars* sc, cudaStream_t st) {
    const long long n_pad = rows_pad + cols_pad;
    CUtensorMap map_hi, map_lo, map_p2, map_h8;
    if (make_plane_map(&map_hi, hi, n_pad, k_pad) || make_plane_map(&map_lo, lo, n_pad, k_pad)) return 1;
    if (fp8lo) {
        const unsigned char* l8 = reinterpret_cast<const unsigned char*>(p2);
        if (make_byte_plane_map(&map_p2, l8, n_pad, k_pad) || make_byte_plane_map(&map_h8, l8 + n_pad * k_pad, n_pad, k_pad)) return 1;
    } else {
        if (make_plane_map(&map_p2, p2, n_pad, k_pad)) return 1;
        map_h8 = map_p2;
    }
    const int2* d_tiles = nullptr;
    int ntiles = 0;
    if (get_cross_tiles(rows_pad / v2::TM, cols_pad / v2::TN, st, &d_tiles, &ntiles)) return 1;
    const Knobs kn = knobs();
    SyrkParams p{};
    p.tiles = d_tiles;
    p.ntiles = ntiles;
    p.num_kb = (int)(k_pad / BK);
    p.K = out;
    p.n = n_rows;
    p.n_cols = n_cols;
    p.row0 = cols_pad;
    p.ldk = ldo;
    p.accumulate = accumulate;
    p.sc = sc;
    p.out_scale = 1.0f;
    p.compact = 0;
    p.fp8lo = fp8lo;
    p.run_kb_fast = kn.run_kb_fast;
    p.run_kb_off = 0;               // a train x test product may pair an individual with itself (positive sums anywhere): short runs for every tile
    p.red_add = kn.red_add;
    p.dbg = kn.dbg;
    CUtensorMap map_out;
    int out_ok = 0;
    make_out_map(&map_out, out, n_rows, n_cols, ldo, &out_ok);
    p.tma_out = (kn.tma_out && out_ok) ? 1 : 0;
    PSTB_CUDA(cudaFuncSetAttribute(v2::k_syrk2, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SYRK2_SMEM));
    if (ntiles < 1) return 0;
    int clusters = sm_count_cached() / 2;
    if (clusters > ntiles) clusters = ntiles;
    if (kn.dynamic && next_counter(st, &p.counter)) return 1;
    v2::k_syrk2<<<2 * clusters, SYRK_THREADS, v2::SYRK2_SMEM, st>>>(map_hi, map_lo, map_p2, map_h8, map_out, p);
    PSTB_AFTER_LAUNCH("k_syrk2");
    return 0;
}

}  // namespace
}  // namespace pstb

using namespace pstb;

extern "C" int64_t pstb_kernel_workspace_bytes(int64_t n_iid, int64_t chunk) {
    if (n_iid < 0) n_iid = 0;
    if (chunk < BK) chunk = BK;
    const long long n_pad = round_up(n_iid > 0 ? n_iid : 1, ROW_PAD), k_pad = round_up(chunk, BK);
    return (int64_t)(3 * n_pad * k_pad * 2 + 1024 + sizeof(Scalars) + (n_pad + 2) * sizeof(double));
}

extern "C" int pstb_syrk_planes(const void* d_hi, const void* d_lo, int64_t n, int64_t n_pad, int64_t k_pad, float* d_K,
                                int64_t ldk, int accumulate, float out_scale, void* stream) {
    if (n <= 0) return 0;
    if (!d_hi || !d_lo || !d_K) return fail("NULL pointer");
    if (ldk < n) return fail("ldk < n");
    return launch_syrk((const __half*)d_hi, (const __half*)d_lo, n, n_pad, k_pad, d_K, ldk, accumulate, nullptr, out_scale,
                       reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int pstb_mirror_lower(float* d_K, int64_t n, int64_t ldk, void* stream) {
    if (n <= 0) return 0;
    if (!d_K) return fail("NULL pointer");
    const unsigned g = (unsigned)((n + 31) / 32);
    k_mirror<<<dim3(g, g), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_K, n, ldk);
    PSTB_AFTER_LAUNCH("k_mirror");
    return 0;
}

extern "C" int pstb_convert_kernel(const float* d_K, int64_t n, void* d_out, int dtype, double scale, void* stream) {
    if (n <= 0) return 0;
    return pstb::convert_range(d_K, (long long)n * n, d_out, dtype, scale, stream);
}

// float32 -> float32 / float64 copy of `total` contiguous entries (a row band of K for the host-buffer kernel entry point)
int pstb::convert_range(const float* d_K, long long total, void* d_out, int dtype, double scale, void* stream) {
    if (total <= 0) return 0;
    if (!d_K || !d_out) return fail("NULL pointer");
    long long grid = (total + 255) / 256;
    const long long cap = (long long)sm_count_cached() * 16;
    if (grid > cap) grid = cap;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == PSTB_F64) k_convert<double><<<(unsigned)grid, 256, 0, st>>>(d_K, total, (double*)d_out, scale);
    else if (dtype == PSTB_F32) k_convert<float><<<(unsigned)grid, 256, 0, st>>>(d_K, total, (float*)d_out, scale);
    else return fail("kernel dtype must be float32 or float64");
    PSTB_AFTER_LAUNCH("k_convert");
    return 0;
}

// phase: bit 0 = first call of a streamed sequence (clear the rank-one accumulators of the exact-dosage path), bit 1 = last call
// (add the rank-one part to K).  A plain call sets both; pstb_snp_kernel_host streams slices through the same workspace and
// applies the rank-one part once.
static std::atomic<int> g_low_term{[] {
    const char* e = getenv("PSTB_SYRK_FP8LO");                  // default for pstb_set_syrk_low_term: 0 = fp16, 1 = fp8, unset = auto
    return e ? (atoi(e) != 0 ? PSTB_LOW_TERM_FP8 : PSTB_LOW_TERM_FP16) : PSTB_LOW_TERM_AUTO;
}()};

extern "C" int pstb_set_syrk_low_term(int mode) {
    if (mode != PSTB_LOW_TERM_FP16 && mode != PSTB_LOW_TERM_FP8 && mode != PSTB_LOW_TERM_AUTO) return g_low_term.load();
    return g_low_term.exchange(mode);
}

// 1 = low term on the fp8 pipe.  `low_term` is the per-call argument (PSTB_LOW_TERM_DEFAULT = the process-wide default above).
// AUTO: the e4m3 rounding error of the low term averages out over the SNPs of the kernel -- relative Frobenius error
// ~5e-6 * sqrt(N / (M_eff + N)) (scripts/emulate_masked_split.py) -- so it is taken when the kernel multiplies at least as many
// SNPs as individuals (Beta weights concentrate on the rare SNPs: 4 x as many) and at least 256.
static int resolve_fp8lo(int low_term, int64_t snps, int64_t n_iid, int mode) {
    int lt = (low_term == PSTB_LOW_TERM_FP16 || low_term == PSTB_LOW_TERM_FP8 || low_term == PSTB_LOW_TERM_AUTO) ? low_term : g_low_term.load();
    if (lt == PSTB_LOW_TERM_AUTO) return (snps >= n_iid * (mode == PSTB_STD_BETA ? 4 : 1) && snps >= 256) ? 1 : 0;
    return lt == PSTB_LOW_TERM_FP8 ? 1 : 0;
}

static int snp_kernel_impl(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid,
                           pstb_axis sid, int count_a1, int mode, double a, double b, int use_stats, double* d_stats,
                           float* d_K, int accumulate, int mirror, void* d_work, int64_t work_bytes, int64_t chunk, int low_term,
                           void* stream, int rank, int world, int compact, int phase = 3, int64_t total_sid = -1,
                           long long tile_begin = 0, long long tile_end = -1, int band_flags = 1, int reserve_sms = 0) {
    // banded mode (tile_end >= 0; pstb_snp_kernel_tiles_band): ONE chunk of SNPs, tiles [tile_begin, tile_end) only.  band_flags bit 0:
    // build the operand planes (first band of the chunk); without it the planes a previous band call left in d_work are reused.
    const bool banded = tile_end >= 0;
    // compact storage only: leave the rank-one vector v in the workspace (pstb_kernel_workspace_rank1) instead of adding it to the tiles;
    // the caller sums / all-reduces the vectors and hands the total to pstb_kernel_from_tiles_range
    const bool defer_rank1 = compact && (((accumulate & 2) != 0) || (banded && (band_flags & 2)));
    accumulate &= 1;
    if (banded && sid.n > chunk) return fail("a band call multiplies one chunk of SNPs (%lld > %lld)", (long long)sid.n, (long long)chunk);
    if (mode != PSTB_STD_UNIT && mode != PSTB_STD_BETA) return fail("kernel needs PSTB_STD_UNIT or PSTB_STD_BETA");
    if (mode == PSTB_STD_BETA && !(a > 0.0 && b > 0.0)) return fail("Beta parameters must be positive");
    if (iid.n < 0 || sid.n < 0) return fail("negative selection length");
    if (iid.n == 0) return 0;
    if (sid.n > 0 && !d_stats) return fail("d_stats is NULL");
    if (compact && pstb_kernel_tile_count(iid.n, rank, world) == 0) {
        // more ranks than tiles: this rank owns nothing, but its caller still gets the per-SNP statistics
        if (sid.n > 0 && !use_stats)
            return read_impl_ex(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, 0, d_stats, nullptr, PSTB_F32, PSTB_ORDER_F,
                                stream, nullptr);
        return 0;
    }
    if (!d_K) return fail("d_K is NULL");
    if (chunk < BK || chunk % BK) return fail("chunk must be a positive multiple of %d", BK);
    if (work_bytes < pstb_kernel_workspace_bytes(iid.n, chunk) || !d_work) return fail("workspace too small (pstb_kernel_workspace_bytes)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long n = iid.n, n_pad = round_up(n, ROW_PAD);
    if (sid.n == 0) {
        if (!accumulate && !compact) PSTB_CUDA(cudaMemsetAsync(d_K, 0, (size_t)n * n * sizeof(float), st));
        return 0;
    }
    const double lnB = (mode == PSTB_STD_BETA) ? lgamma(a) + lgamma(b) - lgamma(a + b) : 0.0;
    const long long k_cap = round_up(chunk, BK);
    char* w = reinterpret_cast<char*>(d_work);
    w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(w) + 1023) & ~(uintptr_t)1023);
    __half* hi = reinterpret_cast<__half*>(w);
    __half* lo = hi + n_pad * k_cap;
    __half* p2 = lo + n_pad * k_cap;
    Scalars* sc = reinterpret_cast<Scalars*>(p2 + n_pad * k_cap);
    double* u = reinterpret_cast<double*>(sc + 1);
    const int dense = (iid.idx == nullptr && iid.step == 1 && (iid.start % 4) == 0) ? 1 : 0;
    // every chunk takes the 2-term exact-dosage GEMM -- missing genotypes and trained statistics included -- unless a trained
    // mean lies outside [0, 2] (k_absmax raises the flag on the device) or PSTB_SYRK_V1=1 / PSTB_SYRK_3TERM=1 (A/B runs) force
    // the 3-term hi/lo split (the 1-CTA kernel has no 2-term path)
    const bool force_slow = (getenv("PSTB_SYRK_V1") && atoi(getenv("PSTB_SYRK_V1")) != 0) ||
                            (getenv("PSTB_SYRK_3TERM") && atoi(getenv("PSTB_SYRK_3TERM")) != 0);
    // (a streamed call passes the SNP count of the whole kernel in total_sid)
    const int fp8lo = force_slow ? 0 : resolve_fp8lo(low_term, total_sid >= 0 ? total_sid : sid.n, iid.n, mode);
    const bool build_planes = !banded || (band_flags & 1);
    if ((phase & 1) && build_planes) PSTB_CUDA(cudaMemsetAsync(u, 0, (size_t)(n_pad + 2) * sizeof(double), st));
    const int2* d_tiles = nullptr;
    int ntiles = 0;
    for (long long c0 = 0; c0 < sid.n; c0 += chunk) {
        const long long ns = (c0 + chunk <= sid.n) ? chunk : sid.n - c0;
        const long long k_pad = round_up(ns, BK);
        if (!build_planes) {                               // later band of the same chunk: planes, scalars and u are in the workspace
            int rc = launch_syrk(hi, lo, n, n_pad, k_pad, d_K, n, accumulate ? 1 : 0, sc, 1.0f, st, rank, world, compact, p2, &d_tiles, &ntiles, fp8lo,
                                 tile_begin, tile_end, reserve_sms);
            if (rc) return rc;
            continue;
        }
        pstb_axis sub = sid;
        sub.n = ns;
        if (sid.idx) sub.idx = sid.idx + c0; else sub.start = sid.start + c0 * sid.step;
        double* st_chunk = d_stats + 2 * c0;
        PSTB_CUDA(cudaMemsetAsync(sc, 0, sizeof(Scalars), st));
        if (force_slow) PSTB_CUDA(cudaMemsetAsync(&sc->need_3term, 0xFF, sizeof(unsigned int), st));
        if (!use_stats) {
            int rc = read_impl_ex(d_packed, ld, iid_count, sid_count, iid, sub, count_a1, mode, a, b, 0, st_chunk, nullptr, PSTB_F32,
                                  PSTB_ORDER_F, stream, nullptr);
            if (rc) return rc;
        }
        k_absmax<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(st_chunk, ns, mode, a, b, lnB, sc);
        PSTB_AFTER_LAUNCH("k_absmax");
        PlaneParams pp{};
        pp.packed = d_packed;
        pp.ld = ld;
        pp.iid_count = iid_count;
        pp.sid_count = sid_count;
        pp.iid = to_axis(iid);
        pp.sid = to_axis(sub);
        pp.count_a1 = count_a1 ? 1 : 0;
        pp.mode = mode;
        pp.a = a;
        pp.b = b;
        pp.lnB = lnB;
        pp.stats = st_chunk;
        pp.sc = sc;
        pp.hi = hi;
        pp.lo = lo;
        pp.p2 = p2;
        pp.fp8lo = fp8lo;
        pp.u = u;
        pp.n_pad = n_pad;
        pp.k_pad = k_pad;
        pp.dense = dense;
        pp.byte_off = dense ? iid.start / 4 : 0;
        const long long ptiles = (k_pad / PT_S) * (n_pad / PT_I);
        launch_planes(pp, ptiles, st);
        PSTB_AFTER_LAUNCH("k_planes");
        int rc = launch_syrk(hi, lo, n, n_pad, k_pad, d_K, n, (accumulate || c0 > 0) ? 1 : 0, sc, 1.0f, st, rank, world, compact, p2,
                             &d_tiles, &ntiles, pp.fp8lo, tile_begin, tile_end, reserve_sms);
        if (rc) return rc;
    }
    if (!force_slow && (phase & 2) && !defer_rank1) {
        if (compact) {
            if (ntiles > 0) {
                // (banded: d_tiles / ntiles are the band's, so is the tile storage)
                k_apply_rank1_tiles<<<(unsigned)ntiles, 256, 0, st>>>(banded ? d_K + tile_begin * 65536LL : d_K, d_tiles, n, u);
                PSTB_AFTER_LAUNCH("k_apply_rank1_tiles");
            }
        } else {
            k_apply_rank1<<<dim3((unsigned)n, (unsigned)((n + 2047) / 2048 > 64 ? 64 : (n + 2047) / 2048)), 256, 0, st>>>(d_K, n, n, u);
            PSTB_AFTER_LAUNCH("k_apply_rank1");
        }
    }
    if (mirror && !compact) return pstb_mirror_lower(d_K, n, n, stream);
    return 0;
}

extern "C" int pstb_snp_kernel(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid,
                               pstb_axis sid, int count_a1, int mode, double a, double b, int use_stats, double* d_stats,
                               float* d_K, int accumulate, int mirror, void* d_work, int64_t work_bytes, int64_t chunk,
                               int low_term, void* stream) {
    return snp_kernel_impl(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, use_stats, d_stats, d_K, accumulate,
                           mirror, d_work, work_bytes, chunk, low_term, stream, 0, 1, 0);
}

int pstb::snp_kernel_slice(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                           int count_a1, int mode, double a, double b, int use_stats, double* d_stats, float* d_K, int accumulate,
                           void* d_work, int64_t work_bytes, int64_t chunk, int low_term, void* stream, int phase, int64_t total_sid) {
    return snp_kernel_impl(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, use_stats, d_stats, d_K, accumulate, 0,
                           d_work, work_bytes, chunk, low_term, stream, 0, 1, 0, phase, total_sid);
}

// ---- K-tile sharding (cfg5: N = 500 000, K = 1 TB does not fit one GPU) -----------------------------------------
extern "C" int64_t pstb_kernel_tile_count(int64_t n_iid, int rank, int world) {
    if (n_iid <= 0 || world < 1 || rank < 0 || rank >= world) return 0;
    std::vector<int2> tiles;
    owned_tiles(n_iid, 2, rank, world, tiles);
    return (int64_t)tiles.size();
}

extern "C" int pstb_kernel_tile_coords(int64_t n_iid, int rank, int world, int32_t* h_ij) {
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world");
    if (!h_ij) return fail("h_ij is NULL");
    std::vector<int2> tiles;
    if (n_iid > 0) owned_tiles(n_iid, 2, rank, world, tiles);
    for (size_t t = 0; t < tiles.size(); ++t) { h_ij[2 * t] = tiles[t].x; h_ij[2 * t + 1] = tiles[t].y; }
    return 0;
}

extern "C" int pstb_snp_kernel_tiles(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid,
                                     pstb_axis sid, int count_a1, int mode, double a, double b, int use_stats, double* d_stats,
                                     float* d_tiles, int rank, int world, int accumulate, void* d_work, int64_t work_bytes,
                                     int64_t chunk, int low_term, void* stream) {
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world");
    return snp_kernel_impl(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, use_stats, d_stats, d_tiles, accumulate,
                           0, d_work, work_bytes, chunk, low_term, stream, rank, world, 1);
}

// what PSTB_LOW_TERM_DEFAULT / _AUTO mean for a kernel of `total_sid` SNPs: callers that split one kernel over several calls (band
// calls, streamed slices) resolve the mode ONCE and pass the explicit value to every call
extern "C" int pstb_resolve_low_term(int low_term, int64_t total_sid, int64_t n_iid, int mode) {
    return resolve_fp8lo(low_term, total_sid, n_iid, mode) ? PSTB_LOW_TERM_FP8 : PSTB_LOW_TERM_FP16;
}

// One band of a compact-tile kernel: tiles [tile_begin, tile_end) of `rank`'s list for ONE chunk of SNPs (sid.n <= chunk).  The
// SNP-sharded multi-GPU path runs its last chunk band by band so that the NCCL all-reduce of finished bands overlaps the
// multiplication of the later ones (`reserve_sms` SMs are left to the collective).  flags bit 0: first band of the chunk (statistics +
// operand planes are built into d_work); later bands reuse them, so the calls of one chunk must not be interleaved with other
// kernel calls on the same workspace.  The rank-one part of the chunk is added to every band's tiles.
extern "C" int pstb_snp_kernel_tiles_band(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid,
                                          pstb_axis sid, int count_a1, int mode, double a, double b, int use_stats, double* d_stats,
                                          float* d_tiles, int rank, int world, int accumulate, void* d_work, int64_t work_bytes,
                                          int64_t chunk, int low_term, int64_t tile_begin, int64_t tile_end, int flags, int reserve_sms,
                                          void* stream) {
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world");
    if (tile_begin < 0 || tile_end < tile_begin) return fail("bad tile range");
    if (low_term != PSTB_LOW_TERM_FP16 && low_term != PSTB_LOW_TERM_FP8) return fail("a band call needs an explicit low_term (FP16 / FP8): every band of a chunk must use the same planes");
    return snp_kernel_impl(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, use_stats, d_stats, d_tiles, accumulate,
                           0, d_work, work_bytes, chunk, low_term, stream, rank, world, 1, 3, -1, tile_begin, tile_end, flags, reserve_sms);
}

// tiles [tile_begin, tile_end) of `rank`'s compact list -> their places in the full symmetric K (the other entries are untouched)
// the rank-one vector v [n_iid] (float64) a kernel call with the "defer" flag left in its workspace
extern "C" double* pstb_kernel_workspace_rank1(void* d_work, int64_t n_iid, int64_t chunk) {
    if (!d_work || n_iid < 1 || chunk < BK) return nullptr;
    const long long n_pad = round_up(n_iid, ROW_PAD), k_cap = round_up(chunk, BK);
    char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_work) + 1023) & ~(uintptr_t)1023);
    __half* p2_end = reinterpret_cast<__half*>(w) + 3 * n_pad * k_cap;
    return reinterpret_cast<double*>(reinterpret_cast<Scalars*>(p2_end) + 1);
}

extern "C" int pstb_kernel_from_tiles_range(const float* d_tiles, int64_t n_iid, int rank, int world, int64_t tile_begin, int64_t tile_end,
                                            float* d_K, const double* d_u, void* stream) {
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world");
    if (n_iid <= 0 || tile_end <= tile_begin) return 0;
    if (!d_tiles || !d_K) return fail("NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int2* coords = nullptr;
    int ntiles = 0;
    if (get_tiles(n_iid, 2, rank, world, st, &coords, &ntiles)) return 1;
    if (tile_begin < 0 || tile_end > ntiles) return fail("bad tile range [%lld, %lld) of %d", (long long)tile_begin, (long long)tile_end, ntiles);
    k_untile<<<dim3((unsigned)(tile_end - tile_begin), 64), 256, 0, st>>>(d_tiles + tile_begin * 65536LL, coords + tile_begin, n_iid, d_K, n_iid, d_u);
    PSTB_AFTER_LAUNCH("k_untile");
    return 0;
}

extern "C" int pstb_float_kernel(const void* d_val, int dtype, int order, int64_t n_iid, int64_t n_sid, float* d_K, int accumulate,
                                 int mirror, void* d_work, int64_t work_bytes, int64_t chunk, void* stream) {
    if (n_iid < 0 || n_sid < 0) return fail("negative shape");
    if (n_iid == 0) return 0;
    if (!d_K) return fail("d_K is NULL");
    if (dtype != PSTB_F32 && dtype != PSTB_F64) return fail("float kernel needs float32 or float64 values");
    if (order != PSTB_ORDER_C && order != PSTB_ORDER_F) return fail("bad order");
    if (chunk < BK || chunk % BK) return fail("chunk must be a positive multiple of %d", BK);
    if (work_bytes < pstb_kernel_workspace_bytes(n_iid, chunk) || !d_work) return fail("workspace too small (pstb_kernel_workspace_bytes)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long n = n_iid, n_pad = round_up(n, ROW_PAD);
    if (n_sid == 0) {
        if (!accumulate) PSTB_CUDA(cudaMemsetAsync(d_K, 0, (size_t)n * n * sizeof(float), st));
        return 0;
    }
    if (!d_val) return fail("d_val is NULL");
    const long long k_cap = round_up(chunk, BK);
    char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_work) + 1023) & ~(uintptr_t)1023);
    __half* hi = reinterpret_cast<__half*>(w);
    __half* lo = hi + n_pad * k_cap;
    Scalars* sc = reinterpret_cast<Scalars*>(lo + n_pad * k_cap);
    const long long si = (order == PSTB_ORDER_C) ? n_sid : 1, sj = (order == PSTB_ORDER_C) ? 1 : n_iid;
    const long long total = n * n_sid;
    PSTB_CUDA(cudaMemsetAsync(sc, 0, sizeof(Scalars), st));
    PSTB_CUDA(cudaMemsetAsync(&sc->need_3term, 0xFF, sizeof(unsigned int), st));       // arbitrary floats: always the 3-term split
    long long g = (total + 255) / 256;
    if (g > (long long)sm_count_cached() * 16) g = (long long)sm_count_cached() * 16;
    if (dtype == PSTB_F32) k_absmax_float<float><<<(unsigned)g, 256, 0, st>>>((const float*)d_val, total, sc);
    else k_absmax_float<double><<<(unsigned)g, 256, 0, st>>>((const double*)d_val, total, sc);
    PSTB_AFTER_LAUNCH("k_absmax_float");
    for (long long c0 = 0; c0 < n_sid; c0 += chunk) {
        const long long ns = (c0 + chunk <= n_sid) ? chunk : n_sid - c0;
        const long long k_pad = round_up(ns, BK);
        dim3 grid((unsigned)(k_pad / 32), (unsigned)(n_pad / 32));
        if (dtype == PSTB_F32) k_split_planes<float><<<grid, 256, 0, st>>>((const float*)d_val, si, sj, n, c0, ns, sc, hi, lo, n_pad, k_pad);
        else k_split_planes<double><<<grid, 256, 0, st>>>((const double*)d_val, si, sj, n, c0, ns, sc, hi, lo, n_pad, k_pad);
        PSTB_AFTER_LAUNCH("k_split_planes");
        int rc = launch_syrk(hi, lo, n, n_pad, k_pad, d_K, n, (accumulate || c0 > 0) ? 1 : 0, sc, 1.0f, st);
        if (rc) return rc;
    }
    if (mirror) return pstb_mirror_lower(d_K, n, n, stream);
    return 0;
}

// ---- train x test kernel: out[i, k] = sum_j x_ij y_kj, both sides standardized with the statistics of the ROW (train) side ----
extern "C" int64_t pstb_cross_kernel_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t chunk) {
    if (n_rows < 0) n_rows = 0;
    if (n_cols < 0) n_cols = 0;
    return pstb_kernel_workspace_bytes(round_up(n_rows > 0 ? n_rows : 1, ROW_PAD) + round_up(n_cols > 0 ? n_cols : 1, ROW_PAD), chunk);
}

extern "C" int pstb_snp_cross_kernel(const uint8_t* d_packed_r, int64_t ld_r, int64_t iid_count_r, int64_t sid_count_r, pstb_axis iid_r,
                                     pstb_axis sid_r, int count_a1_r, const uint8_t* d_packed_c, int64_t ld_c, int64_t iid_count_c,
                                     int64_t sid_count_c, pstb_axis iid_c, pstb_axis sid_c, int count_a1_c, int mode, double a, double b,
                                     int use_stats, double* d_stats, float* d_out, int accumulate, void* d_work, int64_t work_bytes,
                                     int64_t chunk, int low_term, void* stream) {
    if (mode != PSTB_STD_UNIT && mode != PSTB_STD_BETA) return fail("kernel needs PSTB_STD_UNIT or PSTB_STD_BETA");
    if (mode == PSTB_STD_BETA && !(a > 0.0 && b > 0.0)) return fail("Beta parameters must be positive");
    if (iid_r.n < 0 || iid_c.n < 0 || sid_r.n < 0 || sid_c.n < 0) return fail("negative selection length");
    if (sid_r.n != sid_c.n) return fail("both sides must select the same number of SNPs (%lld vs %lld)", (long long)sid_r.n, (long long)sid_c.n);
    if (iid_r.n == 0 || iid_c.n == 0) return 0;
    if (!d_out) return fail("d_out is NULL");
    if (sid_r.n > 0 && !d_stats) return fail("d_stats is NULL");
    if (chunk < BK || chunk % BK) return fail("chunk must be a positive multiple of %d", BK);
    if (work_bytes < pstb_cross_kernel_workspace_bytes(iid_r.n, iid_c.n, chunk) || !d_work)
        return fail("workspace too small (pstb_cross_kernel_workspace_bytes)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long nr = iid_r.n, nc = iid_c.n, rows_pad = round_up(nr, ROW_PAD), cols_pad = round_up(nc, ROW_PAD), n_pad = rows_pad + cols_pad;
    if (sid_r.n == 0) {
        if (!accumulate) PSTB_CUDA(cudaMemsetAsync(d_out, 0, (size_t)nr * nc * sizeof(float), st));
        return 0;
    }
    const double lnB = (mode == PSTB_STD_BETA) ? lgamma(a) + lgamma(b) - lgamma(a + b) : 0.0;
    const long long k_cap = round_up(chunk, BK);
    char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_work) + 1023) & ~(uintptr_t)1023);
    __half* hi = reinterpret_cast<__half*>(w);
    __half* lo = hi + n_pad * k_cap;
    __half* p2 = lo + n_pad * k_cap;
    Scalars* sc = reinterpret_cast<Scalars*>(p2 + n_pad * k_cap);
    double* u = reinterpret_cast<double*>(sc + 1);                                        // rank-one vector v of the column side
    const bool force_slow = getenv("PSTB_SYRK_3TERM") && atoi(getenv("PSTB_SYRK_3TERM")) != 0;
    // AUTO means fp16 here: a train x test product has no diagonal, so its Frobenius norm lacks the mass that makes the e4m3 rounding of
    // the low term small in relative terms (measured 1.1e-5 on the rare-variant fixture all_chr.maf0.001.N300, 290 x 10); FP8 on request
    int lt = (low_term == PSTB_LOW_TERM_FP16 || low_term == PSTB_LOW_TERM_FP8 || low_term == PSTB_LOW_TERM_AUTO) ? low_term : g_low_term.load();
    const int fp8lo = (!force_slow && lt == PSTB_LOW_TERM_FP8) ? 1 : 0;
    PSTB_CUDA(cudaMemsetAsync(u, 0, (size_t)(n_pad + 2) * sizeof(double), st));
    for (long long c0 = 0; c0 < sid_r.n; c0 += chunk) {
        const long long ns = (c0 + chunk <= sid_r.n) ? chunk : sid_r.n - c0;
        const long long k_pad = round_up(ns, BK);
        pstb_axis sub_r = sid_r, sub_c = sid_c;
        sub_r.n = sub_c.n = ns;
        if (sid_r.idx) sub_r.idx = sid_r.idx + c0; else sub_r.start = sid_r.start + c0 * sid_r.step;
        if (sid_c.idx) sub_c.idx = sid_c.idx + c0; else sub_c.start = sid_c.start + c0 * sid_c.step;
        double* st_chunk = d_stats + 2 * c0;
        PSTB_CUDA(cudaMemsetAsync(sc, 0, sizeof(Scalars), st));
        if (force_slow) PSTB_CUDA(cudaMemsetAsync(&sc->need_3term, 0xFF, sizeof(unsigned int), st));
        if (!use_stats) {                                                                 // statistics of the row (train) side
            int rc = read_impl_ex(d_packed_r, ld_r, iid_count_r, sid_count_r, iid_r, sub_r, count_a1_r, mode, a, b, 0, st_chunk, nullptr,
                                  PSTB_F32, PSTB_ORDER_F, stream, nullptr);
            if (rc) return rc;
        }
        k_absmax<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(st_chunk, ns, mode, a, b, lnB, sc);
        PSTB_AFTER_LAUNCH("k_absmax");
        // the e4m3 byte planes of the fp8 low term are addressed over the whole stacked matrix: byte plane 0 ((w h)_lo) at p2,
        // byte plane 1 (g - mu'') at p2 + n_pad * k_pad bytes
        for (int side = 0; side < 2; ++side) {                                            // 0: column operand (rows 0..), 1: row operand
            const pstb_axis& iid = side ? iid_r : iid_c;
            PlaneParams pp{};
            pp.packed = side ? d_packed_r : d_packed_c;
            pp.ld = side ? ld_r : ld_c;
            pp.iid_count = side ? iid_count_r : iid_count_c;
            pp.sid_count = side ? sid_count_r : sid_count_c;
            pp.iid = to_axis(iid);
            pp.sid = to_axis(side ? sub_r : sub_c);
            pp.count_a1 = (side ? count_a1_r : count_a1_c) ? 1 : 0;
            pp.mode = mode;
            pp.a = a;
            pp.b = b;
            pp.lnB = lnB;
            pp.stats = st_chunk;
            pp.sc = sc;
            const long long off = side ? cols_pad * k_pad : 0;
            pp.hi = hi + off;
            pp.lo = lo + off;
            pp.p2 = p2;
            pp.p2_row0 = side ? cols_pad : 0;
            pp.p2_rows = n_pad;
            pp.fp8lo = fp8lo;
            pp.u = side ? nullptr : u;                                                    // v_k belongs to the column side
            pp.n_pad = side ? rows_pad : cols_pad;
            pp.k_pad = k_pad;
            pp.dense = (iid.idx == nullptr && iid.step == 1 && (iid.start % 4) == 0) ? 1 : 0;
            pp.byte_off = pp.dense ? iid.start / 4 : 0;
            const long long ptiles = (k_pad / PT_S) * (pp.n_pad / PT_I);
            launch_planes(pp, ptiles, st);
            PSTB_AFTER_LAUNCH("k_planes");
        }
        int rc = launch_cross(hi, lo, p2, fp8lo, nr, rows_pad, nc, cols_pad, k_pad, d_out, nc, (accumulate || c0 > 0) ? 1 : 0, sc, st);
        if (rc) return rc;
    }
    if (!force_slow) {
        k_apply_colvec<<<dim3((unsigned)nr, (unsigned)((nc + 2047) / 2048 > 64 ? 64 : (nc + 2047) / 2048)), 256, 0, st>>>(d_out, nr, nc, nc, u);
        PSTB_AFTER_LAUNCH("k_apply_colvec");
    }
    return 0;
}

// compact tiles of `rank` (pstb_snp_kernel_tiles layout) -> the full symmetric matrix; with world == 1 that is all of K.  The
// SNP-sharded multi-GPU path all-reduces the compact lower triangle (half the bytes of the square matrix) and expands it here.
extern "C" int pstb_kernel_from_tiles(const float* d_tiles, int64_t n_iid, int rank, int world, float* d_K, void* stream) {
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world");
    if (n_iid <= 0) return 0;
    if (!d_tiles || !d_K) return fail("NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int2* coords = nullptr;
    int ntiles = 0;
    if (get_tiles(n_iid, 2, rank, world, st, &coords, &ntiles)) return 1;
    if (ntiles < 1) return 0;
    k_untile<<<dim3((unsigned)ntiles, 64), 256, 0, st>>>(d_tiles, coords, n_iid, d_K, n_iid, nullptr);
    PSTB_AFTER_LAUNCH("k_untile");
    return 0;
}

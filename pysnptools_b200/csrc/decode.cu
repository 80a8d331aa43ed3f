// decode.cu -- K1 (decode) and K2 (fused decode + per-SNP statistics + standardize) for sm_100a.
//
// Replaces bed_reader's read_f32/f64/i8 as called from pysnptools/snpreader/bed.py:337-343 and the
// read -> standardize_f32/f64 sequence (standardizer.py:109-121) without materialising the raw matrix.
//
// F-order output (iid fastest; the reference's default): one "group" of threads owns one SNP record at a
// time -- a warp when the record is small, a whole CTA when it is large.  The record is staged into the
// group's shared memory by a 1-D TMA bulk copy (cp.async.bulk + mbarrier, double buffered so the next
// record is in flight while the current one is written out), optionally re-packed through the iid
// gather, counted with popc (exact integer statistics), turned into a 4-entry value table in float64,
// and streamed out with 128-bit coalesced stores: every output byte is written once, every packed byte
// read once.
// C-order output (sid fastest): statistics come from the same kernel run with no output, then a
// tile-transposing kernel writes 128-byte row segments.
#include "pstb_common.cuh"

namespace pstb {

struct ReadParams {
    const uint8_t* packed;
    long long ld, iid_count, sid_count;
    Axis iid, sid;
    int count_a1, mode, use_stats;
    double a, b, lnB;
    double* stats;
    void* out;
    long long out_ld;       // F order: elements between output columns (= n_iid_out); C order: n_sid_out
    int dense;              // selected iids are start, start+1, ... with start % 16 == 0
    long long byte_off;     // start / 4 when dense
    int bulk_ok;            // records can be fetched with cp.async.bulk (16-byte aligned)
    unsigned copy_bytes;    // bytes fetched per record (multiple of 16 when bulk_ok)
    unsigned rec_bytes;     // ceil(iid_count / 4)
    unsigned raw_stride;    // shared-memory bytes per raw buffer
    unsigned group_smem;    // shared-memory bytes per group
    int nbuf;               // raw buffers per group (1 or 2)
    int direct;             // record too large for shared memory: gather straight from global, in segments
    long long seg_len;      // outputs per segment (multiple of 16) when direct
    int vec_ok;             // output columns are 16-byte aligned
};

template <typename T>
struct Lut4 {
    T c0, c1, c2, c3;  // value of 2-bit code 0..3
    __device__ __forceinline__ T pick(uint32_t code) const {
        T lo = (code & 1u) ? c1 : c0;
        T hi = (code & 1u) ? c3 : c2;
        return (code & 2u) ? hi : lo;
    }
};

template <typename T>
__device__ __forceinline__ T missing_value();
template <>
__device__ __forceinline__ float missing_value<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double missing_value<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <>
__device__ __forceinline__ int8_t missing_value<int8_t>() { return (int8_t)-127; }

// code 00 -> dosage 0 (2 when count_A1), 01 -> missing, 10 -> 1, 11 -> 2 (0 when count_A1)   [SURVEY Appendix A]
template <typename T>
__device__ __forceinline__ Lut4<T> make_code_lut(int mode, int count_a1, double a, double b, double lnB, double mean, double sd) {
    double v0, v1, v2;
    T vm;
    if (mode == PSTB_STD_NONE) {
        v0 = 0.0; v1 = 1.0; v2 = 2.0;
        vm = missing_value<T>();
    } else {
        double f = (mode == PSTB_STD_BETA) ? beta_factor(mean, a, b, lnB) : 0.0;
        v0 = std_value(mode, 0.0, mean, sd, f);
        v1 = std_value(mode, 1.0, mean, sd, f);
        v2 = std_value(mode, 2.0, mean, sd, f);
        vm = from_double<T>(0.0);
    }
    Lut4<T> l;
    l.c0 = from_double<T>(count_a1 ? v2 : v0);
    l.c1 = vm;
    l.c2 = from_double<T>(v1);
    l.c3 = from_double<T>(count_a1 ? v0 : v2);
    return l;
}

__device__ __forceinline__ long long clampll(long long v, long long hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : v); }

// ---- emit one output column from a dense 2-bit record in shared memory ------------------------------
template <typename T>
__device__ __forceinline__ void emit_column(const unsigned char* rec, T* o, long long n_out, const Lut4<T>& lut, int vec_ok, int gid, int gsize);

template <>
__device__ __forceinline__ void emit_column<float>(const unsigned char* rec, float* o, long long n_out, const Lut4<float>& lut, int vec_ok, int gid, int gsize) {
    if (vec_ok) {
        const long long nq = n_out >> 2;
        float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll 4
        for (long long q = gid; q < nq; q += gsize) {
            uint32_t byte = rec[q];
            float4 v;
            v.x = lut.pick(byte & 3u);
            v.y = lut.pick((byte >> 2) & 3u);
            v.z = lut.pick((byte >> 4) & 3u);
            v.w = lut.pick(byte >> 6);
            __stcs(o4 + q, v);
        }
        long long a = (nq << 2) + gid;
        if (a < n_out) __stcs(o + a, lut.pick((rec[a >> 2] >> (2 * (a & 3))) & 3u));
    } else {
        for (long long a = gid; a < n_out; a += gsize) __stcs(o + a, lut.pick((rec[a >> 2] >> (2 * (a & 3))) & 3u));
    }
}

template <>
__device__ __forceinline__ void emit_column<double>(const unsigned char* rec, double* o, long long n_out, const Lut4<double>& lut, int vec_ok, int gid, int gsize) {
    if (vec_ok) {
        const long long nh = n_out >> 1;
        double2* o2 = reinterpret_cast<double2*>(o);
#pragma unroll 4
        for (long long h = gid; h < nh; h += gsize) {
            uint32_t byte = rec[h >> 1];
            uint32_t sh = (uint32_t)(h & 1) * 4u;
            double2 v;
            v.x = lut.pick((byte >> sh) & 3u);
            v.y = lut.pick((byte >> (sh + 2)) & 3u);
            __stcs(o2 + h, v);
        }
        long long a = (nh << 1) + gid;
        if (a < n_out) __stcs(o + a, lut.pick((rec[a >> 2] >> (2 * (a & 3))) & 3u));
    } else {
        for (long long a = gid; a < n_out; a += gsize) __stcs(o + a, lut.pick((rec[a >> 2] >> (2 * (a & 3))) & 3u));
    }
}

template <>
__device__ __forceinline__ void emit_column<int8_t>(const unsigned char* rec, int8_t* o, long long n_out, const Lut4<int8_t>& lut, int vec_ok, int gid, int gsize) {
    if (vec_ok) {
        const long long nw = n_out >> 4;
        const uint32_t* rec32 = reinterpret_cast<const uint32_t*>(rec);
        uint4* o16 = reinterpret_cast<uint4*>(o);
        for (long long w = gid; w < nw; w += gsize) {
            uint32_t word = rec32[w];
            uint32_t r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t byte = (word >> (8 * k)) & 0xffu;
                uint32_t x0 = (uint8_t)lut.pick(byte & 3u), x1 = (uint8_t)lut.pick((byte >> 2) & 3u);
                uint32_t x2 = (uint8_t)lut.pick((byte >> 4) & 3u), x3 = (uint8_t)lut.pick(byte >> 6);
                r[k] = x0 | (x1 << 8) | (x2 << 16) | (x3 << 24);
            }
            __stcs(o16 + w, make_uint4(r[0], r[1], r[2], r[3]));
        }
        for (long long a = (nw << 4) + gid; a < n_out; a += gsize) o[a] = lut.pick((rec[a >> 2] >> (2 * (a & 3))) & 3u);
    } else {
        for (long long a = gid; a < n_out; a += gsize) o[a] = lut.pick((rec[a >> 2] >> (2 * (a & 3))) & 3u);
    }
}

// ---- K1/K2, F order -------------------------------------------------------------------------------
template <typename T, bool kCta>
__global__ void __launch_bounds__(kCta ? 512 : 256) k_read_f(const ReadParams p) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned int red[3][16];

    const int gsize = kCta ? (int)blockDim.x : 32;
    const int gid = kCta ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    const int g_in_cta = kCta ? 0 : (int)(threadIdx.x >> 5);
    const int groups_per_cta = kCta ? 1 : (int)(blockDim.x >> 5);
    auto gsync = [&]() {
        if (kCta) __syncthreads(); else __syncwarp();
    };

    unsigned char* gs = smem_dyn + (size_t)g_in_cta * p.group_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gs);
    unsigned char* raw0 = gs + 16;
    unsigned char* dense = raw0 + (size_t)p.nbuf * p.raw_stride;

    const long long ngroups = (long long)gridDim.x * groups_per_cta;
    const long long n_out = p.iid.n;
    long long b = (long long)blockIdx.x * groups_per_cta + g_in_cta;

    if (p.bulk_ok) {
        if (gid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_mbar_init();
        }
        gsync();
    }
    auto issue = [&](long long bb, int buf) {
        long long j = clampll(p.sid.at(bb), p.sid_count);
        mbar_expect_tx(&bars[buf], p.copy_bytes);
        bulk_g2s(raw0 + (size_t)buf * p.raw_stride, p.packed + j * p.ld, p.copy_bytes, &bars[buf]);
    };
    if (p.bulk_ok && !p.direct && gid == 0 && b < p.sid.n) issue(b, 0);

    for (uint32_t it = 0; b < p.sid.n; b += ngroups, ++it) {
        const int cur = (p.nbuf == 2) ? (int)(it & 1u) : 0;
        const uint32_t parity = (p.nbuf == 2) ? ((it >> 1) & 1u) : (it & 1u);
        unsigned char* raw = raw0 + (size_t)cur * p.raw_stride;
        const long long nb = b + ngroups;
        const long long j = clampll(p.sid.at(b), p.sid_count);
        const uint8_t* src = p.packed + j * p.ld;
        if (p.direct) {
            // nothing staged
        } else if (p.bulk_ok) {
            if (p.nbuf == 2 && gid == 0 && nb < p.sid.n) issue(nb, cur ^ 1);
            mbar_wait(&bars[cur], parity);
        } else {
            for (unsigned i = gid; i < p.rec_bytes; i += gsize) raw[i] = __ldg(src + i);
            gsync();
        }

        // dense 2-bit record of outputs [seg0, seg0 + seg_n) in shared memory
        auto prepare = [&](long long seg0, long long seg_n) -> const unsigned char* {
            if (!p.direct && p.dense) return raw + p.byte_off;
            const long long nbytes = (seg_n + 3) >> 2;
            for (long long q = gid; q < nbytes; q += gsize) {
                uint32_t byte = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    long long a = (q << 2) + t;
                    if (a < seg_n) {
                        long long i = p.dense ? p.iid.start + seg0 + a : clampll(p.iid.at(seg0 + a), p.iid_count);
                        uint32_t rb = p.direct ? (uint32_t)__ldg(src + (i >> 2)) : (uint32_t)raw[i >> 2];
                        byte |= ((rb >> (2 * (i & 3))) & 3u) << (2 * t);
                    }
                }
                dense[q] = (unsigned char)byte;
            }
            gsync();
            return dense;
        };
        auto count = [&](const unsigned char* rec, long long seg_n, unsigned int& c1, unsigned int& c2, unsigned int& c3) {
            // exact dosage counts with popc over 16 genotypes per word
            const uint32_t* rec32 = reinterpret_cast<const uint32_t*>(rec);
            const long long nwords = (seg_n + 15) >> 4;
            for (long long w = gid; w < nwords; w += gsize) {
                uint32_t word = rec32[w];
                if (w == nwords - 1) {
                    unsigned rem = (unsigned)(seg_n & 15);
                    if (rem) word &= (1u << (2 * rem)) - 1u;
                }
                uint32_t lo = word & 0x55555555u, hi = (word >> 1) & 0x55555555u;
                c1 += __popc(lo & ~hi);
                c2 += __popc(hi & ~lo);
                c3 += __popc(hi & lo);
            }
        };
        const long long nseg = p.direct ? (n_out + p.seg_len - 1) / p.seg_len : 1;
        const unsigned char* rec = nullptr;
        if (nseg == 1) rec = prepare(0, n_out);

        double mean = 0.0, sd = 1.0;
        if (p.mode != PSTB_STD_NONE) {
            if (p.use_stats) {
                mean = p.stats[2 * b];
                sd = p.stats[2 * b + 1];
            } else {
                unsigned int c1 = 0, c2 = 0, c3 = 0;
                if (nseg == 1) {
                    count(rec, n_out, c1, c2, c3);
                } else {
                    for (long long sgi = 0; sgi < nseg; ++sgi) {
                        const long long seg0 = sgi * p.seg_len, seg_n = min(p.seg_len, n_out - seg0);
                        count(prepare(seg0, seg_n), seg_n, c1, c2, c3);
                        gsync();
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
                    c3 += __shfl_xor_sync(0xffffffffu, c3, o);
                }
                if (kCta) {
                    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
                    __syncthreads();
                    if ((threadIdx.x & 31) == 0) { red[0][w] = c1; red[1][w] = c2; red[2][w] = c3; }
                    __syncthreads();
                    c1 = c2 = c3 = 0;
                    for (int k = 0; k < nw; ++k) { c1 += red[0][k]; c2 += red[1][k]; c3 += red[2][k]; }
                }
                const long long c0 = n_out - (long long)c1 - (long long)c2 - (long long)c3;
                stats_from_counts(p.count_a1 ? (long long)c3 : c0, (long long)c2, p.count_a1 ? c0 : (long long)c3, mean, sd);
                if (gid == 0 && p.stats) {
                    p.stats[2 * b] = mean;
                    p.stats[2 * b + 1] = sd;
                }
            }
        }
        if (p.out) {
            const Lut4<T> lut = make_code_lut<T>(p.mode, p.count_a1, p.a, p.b, p.lnB, mean, sd);
            T* o = reinterpret_cast<T*>(p.out) + b * p.out_ld;
            if (nseg == 1) {
                emit_column<T>(rec, o, n_out, lut, p.vec_ok, gid, gsize);
            } else {
                for (long long sgi = 0; sgi < nseg; ++sgi) {
                    const long long seg0 = sgi * p.seg_len, seg_n = min(p.seg_len, n_out - seg0);
                    emit_column<T>(prepare(seg0, seg_n), o + seg0, seg_n, lut, p.vec_ok, gid, gsize);
                    gsync();
                }
            }
        }
        gsync();
        if (p.bulk_ok && !p.direct && p.nbuf == 1 && gid == 0 && nb < p.sid.n) issue(nb, 0);
    }
}

// ---- C order: tile-transposing emit ------------------------------------------------------------------
constexpr int kTileS = 32;    // SNPs per tile (one per lane)
constexpr int kTileI = 512;   // individuals per tile
constexpr int kPitchC = kTileI / 4 + 4;  // bytes; 33 words -> conflict-free column reads

template <typename T>
__global__ void __launch_bounds__(256) k_emit_c(const ReadParams p, long long tiles_i) {
    __shared__ __align__(16) unsigned char codes[kTileS][kPitchC];
    __shared__ T lut_s[kTileS][4];
    const long long tile = blockIdx.x;
    const long long ts = tile / tiles_i, ti = tile % tiles_i;
    const long long b0 = ts * kTileS, i0 = ti * kTileI;
    const long long n_out = p.iid.n;
    const int rows = (int)min((long long)kTileI, n_out - i0);
    const int nsnp = (int)min((long long)kTileS, p.sid.n - b0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // packed bytes of the tile -> shared memory (one byte = 4 consecutive output individuals)
    const int nbytes = (rows + 3) >> 2;
    for (int e = threadIdx.x; e < nsnp * (kTileI / 4); e += blockDim.x) {
        const int s = e / (kTileI / 4), q = e % (kTileI / 4);
        if (q >= nbytes) continue;
        const long long j = clampll(p.sid.at(b0 + s), p.sid_count);
        const uint8_t* src = p.packed + j * p.ld;
        uint32_t byte;
        if (p.dense) {
            byte = __ldg(src + p.byte_off + (i0 >> 2) + q);
        } else {
            byte = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                long long a = i0 + 4 * q + t;
                if (a < n_out) {
                    long long i = clampll(p.iid.at(a), p.iid_count);
                    byte |= ((uint32_t)(__ldg(src + (i >> 2)) >> (2 * (i & 3))) & 3u) << (2 * t);
                }
            }
        }
        codes[s][q] = (unsigned char)byte;
    }
    if (threadIdx.x < nsnp) {
        double mean = 0.0, sd = 1.0;
        if (p.mode != PSTB_STD_NONE) {
            mean = p.stats[2 * (b0 + threadIdx.x)];
            sd = p.stats[2 * (b0 + threadIdx.x) + 1];
        }
        Lut4<T> l = make_code_lut<T>(p.mode, p.count_a1, p.a, p.b, p.lnB, mean, sd);
        lut_s[threadIdx.x][0] = l.c0; lut_s[threadIdx.x][1] = l.c1;
        lut_s[threadIdx.x][2] = l.c2; lut_s[threadIdx.x][3] = l.c3;
    }
    __syncthreads();
    if (lane >= nsnp) return;
    Lut4<T> lut;
    lut.c0 = lut_s[lane][0]; lut.c1 = lut_s[lane][1]; lut.c2 = lut_s[lane][2]; lut.c3 = lut_s[lane][3];
    T* o = reinterpret_cast<T*>(p.out) + i0 * p.out_ld + b0 + lane;
    const int nwarps = blockDim.x >> 5;
#pragma unroll 4
    for (int r = warp; r < rows; r += nwarps) {
        uint32_t code = ((uint32_t)codes[lane][r >> 2] >> (2 * (r & 3))) & 3u;
        __stcs(o + (long long)r * p.out_ld, lut.pick(code));
    }
}

// ---- host side ---------------------------------------------------------------------------------------
static int check_axis(const pstb_axis& ax, int64_t count, const char* name) {
    if (ax.n < 0) return fail("%s.n is negative", name);
    if (ax.idx == nullptr && ax.n > 0) {
        int64_t last = ax.start + (ax.n - 1) * ax.step;
        if (ax.start < 0 || ax.start >= count || last < 0 || last >= count)
            return fail("%s selection [%lld : +%lld*%lld] outside [0, %lld)", name, (long long)ax.start, (long long)ax.n,
                        (long long)ax.step, (long long)count);
    }
    return 0;
}

template <typename T>
static int launch_read(const ReadParams& base, int order, cudaStream_t st) {
    ReadParams p = base;
    const long long n_out = p.iid.n;
    const int sms = sm_count_cached();
    if (order == PSTB_ORDER_C) {
        // statistics first (same fused kernel, no output), then the transposing emit
        if (p.mode != PSTB_STD_NONE && !p.use_stats) {
            ReadParams ps = p;
            ps.out = nullptr;
            int rc = launch_read<T>(ps, PSTB_ORDER_F, st);
            if (rc) return rc;
        }
        if (!p.out) return 0;
        p.out_ld = p.sid.n;
        const long long tiles_i = (n_out + kTileI - 1) / kTileI, tiles_s = (p.sid.n + kTileS - 1) / kTileS;
        const long long tiles = tiles_i * tiles_s;
        if (tiles > 0x7fffffffLL) return fail("C-order read too large for one launch (%lld tiles)", tiles);
        k_emit_c<T><<<(unsigned)tiles, 256, 0, st>>>(p, tiles_i);
        PSTB_AFTER_LAUNCH("k_emit_c");
        return 0;
    }
    p.out_ld = n_out;
    p.vec_ok = p.out && ((reinterpret_cast<uintptr_t>(p.out) & 15u) == 0) && ((n_out * (long long)sizeof(T)) % 16 == 0);
    const unsigned rec16 = (p.rec_bytes + 15u) & ~15u;
    p.bulk_ok = ((reinterpret_cast<uintptr_t>(p.packed) & 15u) == 0) && (p.ld % 16 == 0) && (rec16 <= p.ld || p.sid_count == 0);
    p.copy_bytes = rec16;
    p.raw_stride = rec16;
    const unsigned dense_bytes = p.dense ? 0u : (unsigned)((((n_out + 3) >> 2) + 15) & ~15LL);
    const unsigned max_smem = 220u * 1024u;
    // warp-per-record when two raw buffers + the gather record stay small, else CTA-per-record
    const unsigned warp_group = 16u + 2u * rec16 + dense_bytes;
    if (warp_group <= 12u * 1024u) {
        p.nbuf = 2;
        p.group_smem = warp_group;
        const int warps = 8;
        const unsigned smem = warps * p.group_smem;
        int ctas_per_sm = (int)(max_smem / (smem + 1024u));
        if (ctas_per_sm > 8) ctas_per_sm = 8;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        long long want = (p.sid.n + warps - 1) / warps;
        long long grid = (long long)sms * ctas_per_sm;
        if (grid > want) grid = want;
        if (grid < 1) grid = 1;
        PSTB_CUDA(cudaFuncSetAttribute(k_read_f<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_read_f<T, false><<<(unsigned)grid, warps * 32, smem, st>>>(p);
        PSTB_AFTER_LAUNCH("k_read_f<warp>");
    } else {
        p.nbuf = (16u + 2u * rec16 + dense_bytes <= max_smem / 2) ? 2 : 1;
        p.group_smem = 16u + (unsigned)p.nbuf * rec16 + dense_bytes;
        if (p.group_smem > max_smem) {
            // record (+ gather buffer) larger than shared memory: gather straight from global in segments
            p.direct = 1;
            p.nbuf = 0;
            p.raw_stride = 0;
            p.seg_len = 256 * 1024;                               // 64 KiB of 2-bit codes per segment
            if (p.seg_len > ((n_out + 15) & ~15LL)) p.seg_len = (n_out + 15) & ~15LL;
            p.group_smem = 16u + (unsigned)(p.seg_len / 4);
        }
        int ctas_per_sm = (int)(max_smem / (p.group_smem + 1024u));
        if (ctas_per_sm > 4) ctas_per_sm = 4;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        long long grid = (long long)sms * ctas_per_sm;
        if (grid > p.sid.n) grid = p.sid.n;
        if (grid < 1) grid = 1;
        PSTB_CUDA(cudaFuncSetAttribute(k_read_f<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.group_smem));
        k_read_f<T, true><<<(unsigned)grid, 512, p.group_smem, st>>>(p);
        PSTB_AFTER_LAUNCH("k_read_f<cta>");
    }
    return 0;
}

int read_impl(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
              int count_a1, int mode, double a, double b, int use_stats, double* d_stats, void* d_out, int dtype, int order,
              void* stream) {
    if (iid_count < 0 || sid_count < 0) return fail("negative iid_count / sid_count");
    const int64_t rec = (iid_count + 3) / 4;
    if (ld < rec) return fail("ld (%lld) smaller than ceil(iid_count/4) (%lld)", (long long)ld, (long long)rec);
    if (rec > 0x7fffffff) return fail("iid_count too large");
    if (check_axis(iid, iid_count, "iid") || check_axis(sid, sid_count, "sid")) return 1;
    if (order != PSTB_ORDER_F && order != PSTB_ORDER_C) return fail("order must be PSTB_ORDER_F or PSTB_ORDER_C");
    if (mode != PSTB_STD_NONE && mode != PSTB_STD_UNIT && mode != PSTB_STD_BETA) return fail("bad standardize mode %d", mode);
    if (mode != PSTB_STD_NONE && dtype == PSTB_I8) return fail("standardize needs a float32 / float64 output");
    if (mode != PSTB_STD_NONE && use_stats && !d_stats) return fail("use_stats needs d_stats");
    if (mode != PSTB_STD_NONE && order == PSTB_ORDER_C && !d_stats) return fail("C-order standardize needs d_stats");
    if (mode == PSTB_STD_BETA && !(a > 0.0 && b > 0.0)) return fail("Beta parameters must be positive");
    if (iid.n == 0 || sid.n == 0) return 0;
    if (!d_packed) return fail("d_packed is NULL");
    if (!d_out && (mode == PSTB_STD_NONE || !d_stats)) return fail("nothing to do: d_out and d_stats are NULL");

    ReadParams p{};
    p.packed = d_packed;
    p.ld = ld;
    p.iid_count = iid_count;
    p.sid_count = sid_count;
    p.iid = to_axis(iid);
    p.sid = to_axis(sid);
    p.count_a1 = count_a1 ? 1 : 0;
    p.mode = mode;
    p.use_stats = (mode != PSTB_STD_NONE && use_stats) ? 1 : 0;
    p.a = a;
    p.b = b;
    p.lnB = (mode == PSTB_STD_BETA) ? lgamma(a) + lgamma(b) - lgamma(a + b) : 0.0;
    p.stats = d_stats;
    p.out = d_out;
    p.dense = (iid.idx == nullptr && iid.step == 1 && (iid.start % 16) == 0) ? 1 : 0;
    p.byte_off = p.dense ? iid.start / 4 : 0;
    p.rec_bytes = (unsigned)rec;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case PSTB_F32: return launch_read<float>(p, order, st);
        case PSTB_F64: return launch_read<double>(p, order, st);
        case PSTB_I8: return launch_read<int8_t>(p, order, st);
        default: return fail("bad dtype code %d", dtype);
    }
}

}  // namespace pstb

extern "C" int pstb_decode(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid,
                           pstb_axis sid, int count_a1, void* d_out, int dtype, int order, void* stream) {
    if (!d_out && iid.n > 0 && sid.n > 0) return pstb::fail("d_out is NULL");
    return pstb::read_impl(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, PSTB_STD_NONE, 0.0, 0.0, 0, nullptr, d_out,
                           dtype, order, stream);
}

extern "C" int pstb_decode_standardize(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count,
                                       pstb_axis iid, pstb_axis sid, int count_a1, int mode, double a, double b,
                                       int use_stats, double* d_stats, void* d_out, int dtype, int order, void* stream) {
    return pstb::read_impl(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, use_stats, d_stats, d_out,
                           dtype, order, stream);
}

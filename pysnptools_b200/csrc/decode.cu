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
#include <cstdlib>
#include "pstb_common.cuh"

#ifndef PSTB_ST_HINT
#define PSTB_ST_HINT ".cs"   // streaming (evict-first) stores: the output is never re-read by this kernel
#endif
#ifndef PSTB_READ_MINB
#define PSTB_READ_MINB 4
#endif
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
    int vec_ok;             // output columns: 2 = 32-byte aligned, 1 = 16-byte aligned, 0 = neither
    unsigned int* miss_flag;    // optional: set to 1 when any selected genotype of any processed SNP is missing (K3 picks its GEMM by it)
    int* counter;               // zeroed per launch: records / batches are handed out through it (atomicAdd), not by a static stride --
                                // SMs do not get equal shares of the memory system, and equal shares of the work left 13 % of the SM time idle
    const uint32_t* sel_mask;   // gather: 2 bits per individual (0b01 = selected), built once per call; word [mask_words] = "index vector has repeats"
    long long mask_words;
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
// vec: 2 = column base and length are 32-byte multiples (256-bit stores, sm_100), 1 = 16-byte, 0 = scalar.
// Every lane writes one 32-byte (or 16-byte) piece per instruction, adjacent lanes adjacent pieces, so each
// warp store covers 1 KiB (512 B) of the column and every 32-byte sector is written exactly once.
__device__ __forceinline__ void st256_f32(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
    asm volatile("st.global" PSTB_ST_HINT ".v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4),
                 "f"(a5), "f"(a6), "f"(a7)
                 : "memory");
}
__device__ __forceinline__ void st256_f64(double* p, double a0, double a1, double a2, double a3) {
    asm volatile("st.global" PSTB_ST_HINT ".v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a0), "d"(a1), "d"(a2), "d"(a3) : "memory");
}
__device__ __forceinline__ void st256_b32(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global" PSTB_ST_HINT ".v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

template <typename T>
__device__ __forceinline__ void emit_scalar(const unsigned char* rec, T* o, long long from, long long n_out, const Lut4<T>& lut, int gid, int gsize) {
    for (long long a = from + gid; a < n_out; a += gsize) __stcs(o + a, lut.pick((rec[a >> 2] >> (2 * (a & 3))) & 3u));
}

template <typename T>
__device__ __forceinline__ void emit_column(const unsigned char* rec, T* o, long long n_out, const Lut4<T>& lut, int vec, int gid, int gsize);

template <>
__device__ __forceinline__ void emit_column<float>(const unsigned char* rec, float* o, long long n_out, const Lut4<float>& lut, int vec, int gid, int gsize) {
    if (vec == 2) {
        const long long nh = n_out >> 3;                          // 8 genotypes = 2 packed bytes = 32 output bytes
        const uint16_t* rec16 = reinterpret_cast<const uint16_t*>(rec);
#pragma unroll 4
        for (long long h = gid; h < nh; h += gsize) {
            const uint32_t two = rec16[h];
            st256_f32(o + (h << 3), lut.pick(two & 3u), lut.pick((two >> 2) & 3u), lut.pick((two >> 4) & 3u), lut.pick((two >> 6) & 3u),
                      lut.pick((two >> 8) & 3u), lut.pick((two >> 10) & 3u), lut.pick((two >> 12) & 3u), lut.pick(two >> 14));
        }
        emit_scalar<float>(rec, o, nh << 3, n_out, lut, gid, gsize);
    } else if (vec == 1) {
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
        emit_scalar<float>(rec, o, nq << 2, n_out, lut, gid, gsize);
    } else {
        emit_scalar<float>(rec, o, 0, n_out, lut, gid, gsize);
    }
}

template <>
__device__ __forceinline__ void emit_column<double>(const unsigned char* rec, double* o, long long n_out, const Lut4<double>& lut, int vec, int gid, int gsize) {
    if (vec == 2) {
        const long long nq = n_out >> 2;                          // 4 genotypes = 1 packed byte = 32 output bytes
#pragma unroll 4
        for (long long q = gid; q < nq; q += gsize) {
            const uint32_t byte = rec[q];
            st256_f64(o + (q << 2), lut.pick(byte & 3u), lut.pick((byte >> 2) & 3u), lut.pick((byte >> 4) & 3u), lut.pick(byte >> 6));
        }
        emit_scalar<double>(rec, o, nq << 2, n_out, lut, gid, gsize);
    } else if (vec == 1) {
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
        emit_scalar<double>(rec, o, nh << 1, n_out, lut, gid, gsize);
    } else {
        emit_scalar<double>(rec, o, 0, n_out, lut, gid, gsize);
    }
}

// int8 output: the four table values fit one 32-bit word, so PRMT (byte permute) looks up all four genotypes of a packed
// byte at once -- the selector nibbles are the 2-bit codes themselves.
__device__ __forceinline__ uint32_t i8_lutword(const Lut4<int8_t>& lut) {
    return (uint32_t)(uint8_t)lut.c0 | ((uint32_t)(uint8_t)lut.c1 << 8) | ((uint32_t)(uint8_t)lut.c2 << 16) | ((uint32_t)(uint8_t)lut.c3 << 24);
}
__device__ __forceinline__ uint32_t i8x4(uint32_t lutword, uint32_t byte) {
    const uint32_t sel = (byte & 0x3u) | ((byte & 0xcu) << 2) | ((byte & 0x30u) << 4) | ((byte & 0xc0u) << 6);
    return __byte_perm(lutword, 0u, sel);
}

template <>
__device__ __forceinline__ void emit_column<int8_t>(const unsigned char* rec, int8_t* o, long long n_out, const Lut4<int8_t>& lut, int vec, int gid, int gsize) {
    const uint32_t lw = i8_lutword(lut);
    if (vec == 2 && ((reinterpret_cast<uintptr_t>(rec) & 7u) == 0)) {
        const long long nw = n_out >> 5;                          // 32 genotypes = 8 packed bytes = 32 output bytes
        const uint2* rec64 = reinterpret_cast<const uint2*>(rec);
        for (long long w = gid; w < nw; w += gsize) {
            const uint2 two = rec64[w];
            uint32_t r[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r[k] = i8x4(lw, (two.x >> (8 * k)) & 0xffu);
                r[4 + k] = i8x4(lw, (two.y >> (8 * k)) & 0xffu);
            }
            st256_b32(o + (w << 5), r);
        }
        emit_scalar<int8_t>(rec, o, nw << 5, n_out, lut, gid, gsize);
    } else if (vec >= 1) {
        const long long nw = n_out >> 4;
        const uint32_t* rec32 = reinterpret_cast<const uint32_t*>(rec);
        uint4* o16 = reinterpret_cast<uint4*>(o);
        for (long long w = gid; w < nw; w += gsize) {
            uint32_t word = rec32[w];
            uint32_t r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = i8x4(lw, (word >> (8 * k)) & 0xffu);
            __stcs(o16 + w, make_uint4(r[0], r[1], r[2], r[3]));
        }
        emit_scalar<int8_t>(rec, o, nw << 4, n_out, lut, gid, gsize);
    } else {
        emit_scalar<int8_t>(rec, o, 0, n_out, lut, gid, gsize);
    }
}

// ---- K1/K2, F order -------------------------------------------------------------------------------
template <typename T, bool kCta>
__global__ void __launch_bounds__(kCta ? 512 : 256, kCta ? 2 : PSTB_READ_MINB) k_read_f(const ReadParams p) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned int red[3][16];
    __shared__ int s_claim;

    const int gsize = kCta ? (int)blockDim.x : 32;
    const int gid = kCta ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    const int g_in_cta = kCta ? 0 : (int)(threadIdx.x >> 5);
    auto gsync = [&]() {
        if (kCta) __syncthreads(); else __syncwarp();
    };

    unsigned char* gs = smem_dyn + (size_t)g_in_cta * p.group_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gs);
    unsigned char* raw0 = gs + 16;
    unsigned char* dense = raw0 + (size_t)p.nbuf * p.raw_stride;

    const long long n_out = p.iid.n;
    // the next record of this group (warp or CTA), claimed from the launch's counter
    auto claim = [&]() -> long long {
        if (kCta) {
            __syncthreads();                                        // the previous claim has been read by everyone
            if (threadIdx.x == 0) s_claim = atomicAdd(p.counter, 1);
            __syncthreads();
            return (long long)s_claim;
        }
        int v = 0;
        if (gid == 0) v = atomicAdd(p.counter, 1);
        return (long long)__shfl_sync(0xffffffffu, v, 0);
    };

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
    long long b = claim(), nb = 0;
    if (p.bulk_ok && !p.direct && gid == 0 && b < p.sid.n) issue(b, 0);

    for (uint32_t it = 0; b < p.sid.n; b = nb, ++it) {
        const int cur = (p.nbuf == 2) ? (int)(it & 1u) : 0;
        const uint32_t parity = (p.nbuf == 2) ? ((it >> 1) & 1u) : (it & 1u);
        unsigned char* raw = raw0 + (size_t)cur * p.raw_stride;
        nb = claim();
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
                if (gid == 0 && c1 && p.miss_flag) *p.miss_flag = 1u;
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

// ---- K1/K2, F order, short records (N <= 4096): a warp takes a batch of R adjacent records ---------------------------------
// k_read_f spends a fixed chain per record (wait for the copy, 15 shuffles, serial fp64 statistics, table) that a 75-250 byte
// record cannot amortise (N = 300: 30 % of the HBM peak, N = 1 000: 61 %).  Here ONE bulk copy stages R records (they are
// adjacent in the store), lane r counts record r on its own and computes its statistics and value table -- R records at once,
// no shuffles -- and the warp then emits the records one after the other with the table broadcast from lane r.
static const long long kSmallLd = [] {                           // records up to this pitch take the batched kernel (A/B knob: PSTB_SMALL_LD)
    const char* e = getenv("PSTB_SMALL_LD");
    const long long v = e ? atoll(e) : 0;
    return (v >= 16 && v <= 4096) ? v : 1024;                  // measured: 1024 gains 1-2.5 points for N = 2 049 .. 4 096; 2048 loses 20 at N = 6 000
}();

template <typename T>
__device__ __forceinline__ T shfl_t(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <>
__device__ __forceinline__ int8_t shfl_t<int8_t>(int8_t v, int src) { return (int8_t)__shfl_sync(0xffffffffu, (int)v, src); }

template <typename T>
__global__ void __launch_bounds__(256) k_read_f_small(const ReadParams p, int R) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned buf_bytes = (unsigned)R * (unsigned)p.ld;
    unsigned char* ws = smem_dyn + (size_t)w * (16u + 2u * buf_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ws);
    unsigned char* raw0 = ws + 16;
    const long long n_out = p.iid.n;
    const long long nbatch = (p.sid.n + R - 1) / R;
    auto claim = [&]() -> long long {
        int v = 0;
        if (lane == 0) v = atomicAdd(p.counter, 1);
        return (long long)__shfl_sync(0xffffffffu, v, 0);
    };
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    auto issue = [&](long long batch, int buf) {
        const long long s0 = batch * R;
        const unsigned bytes = (unsigned)min((long long)R, p.sid.n - s0) * (unsigned)p.ld;
        mbar_expect_tx(&bars[buf], bytes);
        bulk_g2s(raw0 + (size_t)buf * buf_bytes, p.packed + (p.sid.start + s0) * p.ld, bytes, &bars[buf]);
    };
    long long bt = claim(), nbt = 0;
    if (lane == 0 && bt < nbatch) issue(bt, 0);
    const long long nwords = (n_out + 15) >> 4;
    for (uint32_t it = 0; bt < nbatch; bt = nbt, ++it) {
        const int cur = (int)(it & 1u);
        nbt = claim();
        if (lane == 0 && nbt < nbatch) issue(nbt, cur ^ 1);
        mbar_wait(&bars[cur], (it >> 1) & 1u);
        const unsigned char* raw = raw0 + (size_t)cur * buf_bytes + p.byte_off;
        const long long s0 = bt * R;
        const int nr = (int)min((long long)R, p.sid.n - s0);
        double mean = 0.0, sd = 1.0;
        if (p.mode != PSTB_STD_NONE && lane < nr) {
            if (p.use_stats) {
                mean = p.stats[2 * (s0 + lane)];
                sd = p.stats[2 * (s0 + lane) + 1];
            } else {
                const uint32_t* rec32 = reinterpret_cast<const uint32_t*>(raw + (size_t)lane * p.ld);
                unsigned int c1 = 0, c2 = 0, c3 = 0;
                for (long long k = 0; k < nwords; ++k) {
                    uint32_t word = rec32[k];
                    if (k == nwords - 1) {
                        const unsigned rem = (unsigned)(n_out & 15);
                        if (rem) word &= (1u << (2 * rem)) - 1u;
                    }
                    const uint32_t lo = word & 0x55555555u, hi = (word >> 1) & 0x55555555u;
                    c1 += __popc(lo & ~hi);
                    c2 += __popc(hi & ~lo);
                    c3 += __popc(hi & lo);
                }
                const long long c0 = n_out - (long long)c1 - (long long)c2 - (long long)c3;
                stats_from_counts(p.count_a1 ? (long long)c3 : c0, (long long)c2, p.count_a1 ? c0 : (long long)c3, mean, sd);
                if (c1 && p.miss_flag) *p.miss_flag = 1u;
                if (p.stats) {
                    p.stats[2 * (s0 + lane)] = mean;
                    p.stats[2 * (s0 + lane) + 1] = sd;
                }
            }
        }
        if (p.out) {
            const Lut4<T> mine = make_code_lut<T>(p.mode, p.count_a1, p.a, p.b, p.lnB, mean, sd);
            T* o = reinterpret_cast<T*>(p.out) + s0 * p.out_ld;
            for (int r = 0; r < nr; ++r, o += p.out_ld) {
                Lut4<T> lut;
                lut.c0 = shfl_t<T>(mine.c0, r);
                lut.c1 = shfl_t<T>(mine.c1, r);
                lut.c2 = shfl_t<T>(mine.c2, r);
                lut.c3 = shfl_t<T>(mine.c3, r);
                emit_column<T>(raw + (size_t)r * p.ld, o, n_out, lut, p.vec_ok, lane, 32);
            }
        }
        __syncwarp();
    }
}

// ---- statistics only (no output), contiguous individuals --------------------------------------------------
// The first pass of a C-order standardizing read and of every K3 chunk is a pure read: k_read_f's one-record-ahead TMA
// staging is latency-bound there (2.6 TB/s on cfg2-sized records, 1.3 TB/s at 12.5 KB).  Here every warp streams whole
// records straight into registers with 128-bit loads, four loads per lane in flight, and counts with popc.
__global__ void __launch_bounds__(256) k_stats_dense(const ReadParams p, int R) {
    // a warp takes R consecutive records at a time (R a power of two <= 32); lane r keeps the counts of record r, so the fp64
    // mean / sd arithmetic runs once per R records with R lanes busy and the statistics leave as one coalesced store
    const int lane = threadIdx.x & 31;
    const long long n_out = p.iid.n;
    const long long full16 = n_out >> 6;                       // 128-bit words holding 64 selected genotypes each
    const long long tail0 = full16 << 6;                       // first genotype of the tail
    const long long tail_words = (n_out - tail0 + 15) >> 4;    // 32-bit words of the tail (last one masked)
    auto claim = [&]() -> long long {                          // the warp's next block of R records
        int v = 0;
        if (lane == 0) v = atomicAdd(p.counter, 1);
        return (long long)__shfl_sync(0xffffffffu, v, 0) * R;
    };
    for (long long b0 = claim(); b0 < p.sid.n; b0 = claim()) {
        unsigned int k1 = 0, k2 = 0, k3 = 0;
        const int nrec = (int)min((long long)R, p.sid.n - b0);
        for (int r = 0; r < nrec; ++r) {
            const long long j = clampll(p.sid.at(b0 + r), p.sid_count);
            const uint8_t* src = p.packed + j * p.ld + p.byte_off;
            const uint4* src16 = reinterpret_cast<const uint4*>(src);
            // POPC runs at 16 lanes / clock / SM: three per word kept that pipe 75 % busy at 4.6 TB/s.  Two words per POPC instead --
            // the low bits of word A on the even positions and those of word B on the odd ones (likewise the high bits and the
            // "both bits set" flags) -- and the three dosage counts follow from the three sums.
            unsigned int n_lo = 0, n_hi = 0, n_both = 0;
            auto add2 = [&](uint32_t wa, uint32_t wb) {
                const uint32_t m = 0x55555555u, a1 = wa >> 1, b1 = wb << 1;
                n_lo += __popc((wa & m) | (b1 & ~m));
                n_hi += __popc((a1 & m) | (wb & ~m));
                n_both += __popc((wa & a1 & m) | (wb & b1 & ~m));
            };
            long long w = lane;
            for (; w + 96 < full16; w += 128) {
                const uint4 v0 = __ldg(src16 + w), v1 = __ldg(src16 + w + 32), v2 = __ldg(src16 + w + 64), v3 = __ldg(src16 + w + 96);
                add2(v0.x, v0.y); add2(v0.z, v0.w);
                add2(v1.x, v1.y); add2(v1.z, v1.w);
                add2(v2.x, v2.y); add2(v2.z, v2.w);
                add2(v3.x, v3.y); add2(v3.z, v3.w);
            }
            for (; w < full16; w += 32) {
                const uint4 v = __ldg(src16 + w);
                add2(v.x, v.y); add2(v.z, v.w);
            }
            const uint32_t* tail = reinterpret_cast<const uint32_t*>(src + (tail0 >> 2));
            for (long long t = lane; t < tail_words; t += 32) {
                // the last word may reach into the record's ld padding: masked to the selected genotypes
                uint32_t word = __ldg(tail + t);
                const long long left = n_out - tail0 - 16 * t;
                if (left < 16) word &= (1u << (2 * (unsigned)left)) - 1u;
                add2(word, 0u);
            }
            n_lo = __reduce_add_sync(0xffffffffu, n_lo);
            n_hi = __reduce_add_sync(0xffffffffu, n_hi);
            const unsigned int c3 = __reduce_add_sync(0xffffffffu, n_both);
            const unsigned int c1 = n_lo - c3, c2 = n_hi - c3;          // code 01 = low bit only, 10 = high bit only, 11 = both
            if (lane == r) { k1 = c1; k2 = c2; k3 = c3; }
        }
        if (lane < nrec) {
            const long long c0 = n_out - (long long)k1 - (long long)k2 - (long long)k3;
            double mean, sd;
            stats_from_counts(p.count_a1 ? (long long)k3 : c0, (long long)k2, p.count_a1 ? c0 : (long long)k3, mean, sd);
            if (k1 && p.miss_flag) *p.miss_flag = 1u;
            reinterpret_cast<double2*>(p.stats)[b0 + lane] = make_double2(mean, sd);
        }
    }
}

// ---- K1/K2, F order, gathered individuals: S records per batch share one pass over the index vector ---------
// The iid index vector costs 4 bytes per output genotype -- as much as the float32 output itself -- so it is read
// once per batch of S staged records (TMA bulk copies on one mbarrier), every thread turning its 4 indices into
// S re-packed bytes.  The raw records are dead after the re-pack, so the next batch's bulk copies are issued
// before the counts / emit and overlap with the HBM writes.
constexpr int kGatherMaxS = 16;

template <typename T>
__global__ void __launch_bounds__(1024, 1) k_read_f_gather(const ReadParams p, int S, unsigned dense_stride) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned int cnt[kGatherMaxS][3];
    __shared__ double st_s[kGatherMaxS][2];
    __shared__ double lut_d[kGatherMaxS][4];
    __shared__ int s_claim[2];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_dyn);
    unsigned char* raw0 = smem_dyn + 16;
    unsigned char* dense0 = raw0 + (size_t)S * p.raw_stride;
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long n_out = p.iid.n;
    const long long nbatch = (p.sid.n + S - 1) / S;
    const long long nbytes = (n_out + 3) >> 2;
    const bool idx_vec = p.iid.idx && ((reinterpret_cast<uintptr_t>(p.iid.idx) & 15u) == 0);

    auto issue = [&](long long batch) {
        const long long b0 = batch * S;
        const int ns = (int)min((long long)S, p.sid.n - b0);
        mbar_expect_tx(bar, (uint32_t)ns * p.copy_bytes);
        for (int s = 0; s < ns; ++s) {
            long long j = clampll(p.sid.at(b0 + s), p.sid_count);
            bulk_g2s(raw0 + (size_t)s * p.raw_stride, p.packed + j * p.ld, p.copy_bytes, bar);
        }
    };
    // batches are claimed from the launch's counter by thread 0, one ahead (slot it & 1 holds the batch of iteration it)
    if (tid == 0) {
        if (p.bulk_ok) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        s_claim[0] = atomicAdd(p.counter, 1);
        if (p.bulk_ok && s_claim[0] < nbatch) issue(s_claim[0]);
    }
    __syncthreads();

    for (uint32_t it = 0;; ++it) {
        const long long batch = s_claim[it & 1u];
        if (batch >= nbatch) break;
        const long long b0 = batch * S;
        const int ns = (int)min((long long)S, p.sid.n - b0);
        if (p.bulk_ok) {
            mbar_wait(bar, it & 1u);
        } else {
            for (int s = 0; s < ns; ++s) {
                const uint8_t* src = p.packed + clampll(p.sid.at(b0 + s), p.sid_count) * p.ld;
                for (unsigned i = tid; i < p.rec_bytes; i += nt) raw0[(size_t)s * p.raw_stride + i] = __ldg(src + i);
            }
            __syncthreads();
        }
        if (tid < kGatherMaxS * 3) (&cnt[0][0])[tid] = 0;

        for (long long q = tid; q < nbytes; q += nt) {
            uint32_t boff[4], sh[4], keep = 0;
            if (idx_vec && (q << 2) + 3 < n_out) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.iid.idx) + q);
                const uint32_t iv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    uint32_t i = iv[t] < (uint32_t)p.iid_count ? iv[t] : (uint32_t)p.iid_count - 1u;
                    boff[t] = i >> 2;
                    sh[t] = 2u * (i & 3u);
                }
                keep = 0xffu;
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const long long a = (q << 2) + t;
                    long long i = 0;
                    if (a < n_out) {
                        i = clampll(p.iid.at(a), p.iid_count);
                        keep |= 3u << (2 * t);
                    }
                    boff[t] = (uint32_t)(i >> 2);
                    sh[t] = 2u * (uint32_t)(i & 3);
                }
            }
            for (int s = 0; s < ns; ++s) {
                const unsigned char* r = raw0 + (size_t)s * p.raw_stride;
                uint32_t byte = ((uint32_t)(r[boff[0]] >> sh[0]) & 3u) | (((uint32_t)(r[boff[1]] >> sh[1]) & 3u) << 2) |
                                (((uint32_t)(r[boff[2]] >> sh[2]) & 3u) << 4) | (((uint32_t)(r[boff[3]] >> sh[3]) & 3u) << 6);
                dense0[(size_t)s * dense_stride + q] = (unsigned char)(byte & keep);
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int nbt = atomicAdd(p.counter, 1);
            s_claim[(it + 1u) & 1u] = nbt;                           // read after the barrier that ends this iteration
            if (p.bulk_ok && nbt < nbatch) issue(nbt);               // raw buffers are free again
        }

        if (p.mode != PSTB_STD_NONE) {
            if (!p.use_stats) {
                const long long nwords = (n_out + 15) >> 4;
                for (int s = 0; s < ns; ++s) {
                    const uint32_t* rec32 = reinterpret_cast<const uint32_t*>(dense0 + (size_t)s * dense_stride);
                    unsigned int c1 = 0, c2 = 0, c3 = 0;
                    for (long long w = tid; w < nwords; w += nt) {
                        uint32_t word = rec32[w];
                        if (w == nwords - 1) {
                            unsigned rem = (unsigned)(n_out & 15);
                            if (rem) word &= (1u << (2 * rem)) - 1u;
                        }
                        uint32_t lo = word & 0x55555555u, hi = (word >> 1) & 0x55555555u;
                        c1 += __popc(lo & ~hi);
                        c2 += __popc(hi & ~lo);
                        c3 += __popc(hi & lo);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
                        c3 += __shfl_xor_sync(0xffffffffu, c3, o);
                    }
                    if ((tid & 31) == 0) {
                        atomicAdd(&cnt[s][0], c1);
                        atomicAdd(&cnt[s][1], c2);
                        atomicAdd(&cnt[s][2], c3);
                    }
                }
                __syncthreads();
            }
            if (tid < ns) {
                double mean, sd;
                if (p.use_stats) {
                    mean = p.stats[2 * (b0 + tid)];
                    sd = p.stats[2 * (b0 + tid) + 1];
                } else {
                    const long long c1 = cnt[tid][0], c2 = cnt[tid][1], c3 = cnt[tid][2], c0 = n_out - c1 - c2 - c3;
                    stats_from_counts(p.count_a1 ? c3 : c0, c2, p.count_a1 ? c0 : c3, mean, sd);
                    if (c1 && p.miss_flag) *p.miss_flag = 1u;
                    if (p.stats) {
                        p.stats[2 * (b0 + tid)] = mean;
                        p.stats[2 * (b0 + tid) + 1] = sd;
                    }
                }
                st_s[tid][0] = mean;
                st_s[tid][1] = sd;
            }
        }
        if (tid < ns) {
            const Lut4<T> l = make_code_lut<T>(p.mode, p.count_a1, p.a, p.b, p.lnB, st_s[tid][0], st_s[tid][1]);
            lut_d[tid][0] = (double)l.c0; lut_d[tid][1] = (double)l.c1; lut_d[tid][2] = (double)l.c2; lut_d[tid][3] = (double)l.c3;
        }
        __syncthreads();
        if (p.out) {
            for (int s = 0; s < ns; ++s) {
                Lut4<T> lut;
                lut.c0 = (T)lut_d[s][0]; lut.c1 = (T)lut_d[s][1]; lut.c2 = (T)lut_d[s][2]; lut.c3 = (T)lut_d[s][3];
                emit_column<T>(dense0 + (size_t)s * dense_stride, reinterpret_cast<T*>(p.out) + (b0 + s) * p.out_ld, n_out, lut,
                               p.vec_ok, tid, nt);
            }
        }
        __syncthreads();
    }
}

// The selection mask below is a stream-ordered allocation (cudaMallocAsync: safe for concurrent callers on different streams).  The
// default memory pool hands its memory back to the driver at every synchronisation unless told otherwise, which made every gathered
// read pay a driver allocation on the HOST side -- 0.1 to several ms between the caller's events: cfg4 measured anywhere from 4.3 to
// 11.8 ms per call for the same 4.3 ms kernel.  Keep up to 64 MiB in the pool of each device this library allocates from.
static void keep_pool_memory(cudaStream_t) {
    static thread_local int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t cur = 0;
        if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < (64ull << 20)) {
            uint64_t keep = 64ull << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaGetLastError();
    done_for = dev;
}

// Selection mask for the statistics of gathered reads: bit 2*(i%16) of word i/16 is set when individual i is selected.
// Per-SNP dosage counts over the selected individuals then are masked popcounts over the raw record -- a coalesced
// sweep instead of a second random gather.  A repeated index (the counts would need multiplicities) raises the flag in
// the extra last word and the kernels fall back to counting through the gather.
__global__ void k_build_sel_mask(Axis iid, long long iid_count, uint32_t* mask, long long mask_words) {
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < iid.n; a += (long long)gridDim.x * blockDim.x) {
        long long i = iid.at(a);
        i = i < 0 ? 0 : (i >= iid_count ? iid_count - 1 : i);
        const uint32_t bit = 1u << (2 * (i & 15));
        const uint32_t old = atomicOr(mask + (i >> 4), bit);
        if (old & bit) mask[mask_words] = 1u;
    }
}

// ---- gathered individuals, 4 records at a time, byte-interleaved in shared memory ---------------------------------
// The re-pack above is bound by shared-memory wavefronts (one random byte load per output code, ~3.5-way bank
// conflicts).  Here the four records of a batch are interleaved while they are loaded -- word B of the tile holds
// byte B of all four SNPs -- so ONE random 32-bit load serves an individual for four SNPs, and the values are written
// straight from the gather (no dense intermediate): pass 1 counts, pass 2 emits 16 contiguous bytes per lane and SNP.
template <typename T>
__device__ __forceinline__ void store4(T* o, T a, T b, T c, T d, int vec);
template <>
__device__ __forceinline__ void store4<float>(float* o, float a, float b, float c, float d, int vec) {
    if (vec) __stcs(reinterpret_cast<float4*>(o), make_float4(a, b, c, d));
    else { __stcs(o, a); __stcs(o + 1, b); __stcs(o + 2, c); __stcs(o + 3, d); }
}
template <>
__device__ __forceinline__ void store4<double>(double* o, double a, double b, double c, double d, int vec) {
    if (vec == 2) st256_f64(o, a, b, c, d);
    else if (vec == 1) { __stcs(reinterpret_cast<double2*>(o), make_double2(a, b)); __stcs(reinterpret_cast<double2*>(o) + 1, make_double2(c, d)); }
    else { __stcs(o, a); __stcs(o + 1, b); __stcs(o + 2, c); __stcs(o + 3, d); }
}
template <>
__device__ __forceinline__ void store4<int8_t>(int8_t* o, int8_t a, int8_t b, int8_t c, int8_t d, int vec) {
    if (vec) *reinterpret_cast<uint32_t*>(o) = (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24);
    else { o[0] = a; o[1] = b; o[2] = c; o[3] = d; }
}

// kStaged: the 4 raw records of the NEXT batch are fetched by TMA bulk copies into a staging area while the current
// batch is counted and written (reads queue behind the write flood for microseconds; they must never be waited for).
template <typename T, bool kStaged>
__global__ void __launch_bounds__(kStaged ? 1024 : 512, kStaged ? 1 : 2) k_read_f_gather4(const ReadParams p) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t stage_bar;
    __shared__ int s_claim[2];                                     // batch of iteration it in slot it & 1 (claimed one ahead by thread 0)
    __shared__ unsigned int cnt[4][3];
    __shared__ double st_s[4][2];
    __shared__ double lut_d[4][4];
    uint32_t* inter32 = reinterpret_cast<uint32_t*>(smem_dyn);     // inter32[B] = byte B of SNPs 0..3 of the batch
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long n_out = p.iid.n;
    const long long nbatch = (p.sid.n + 3) >> 2;
    const long long nq = (n_out + 3) >> 2;
    const unsigned rec_words = (p.rec_bytes + 3u) >> 2;
    const bool word_ok = kStaged || (((reinterpret_cast<uintptr_t>(p.packed) & 3u) == 0) && (p.ld % 4 == 0));
    const bool idx_vec = p.iid.idx && ((reinterpret_cast<uintptr_t>(p.iid.idx) & 15u) == 0);
    unsigned char* raw0 = smem_dyn + 16u * (size_t)rec_words;     // staging area (kStaged): 4 records of raw_stride bytes
    auto issue = [&](long long batch) {
        const long long bb = batch << 2;
        const int nn = (int)min(4LL, p.sid.n - bb);
        mbar_expect_tx(&stage_bar, 4u * p.copy_bytes);
        for (int s = 0; s < 4; ++s)
            bulk_g2s(raw0 + (size_t)s * p.raw_stride, p.packed + clampll(p.sid.at(bb + min(s, nn - 1)), p.sid_count) * p.ld, p.copy_bytes, &stage_bar);
    };
    if (tid == 0) {
        if (kStaged) {
            mbar_init(&stage_bar, 1);
            fence_mbar_init();
        }
        s_claim[0] = atomicAdd(p.counter, 1);
        if (kStaged && s_claim[0] < nbatch) issue(s_claim[0]);
    }
    __syncthreads();
    uint32_t iter = 0;
    // statistics of gathered reads without a second gather: masked popcounts over the raw records (see k_build_sel_mask)
    const bool mask_count = kStaged && p.mode != PSTB_STD_NONE && !p.use_stats && p.sel_mask != nullptr && __ldg(p.sel_mask + p.mask_words) == 0u;
    if (tid < 12) (&cnt[0][0])[tid] = 0;
    __syncthreads();

    // the 4 iid indices of output quad q (fetched one iteration ahead: the index vector lives in L2) ...
    auto fetch_quad = [&](long long q) -> uint4 {
        if (q >= nq) return make_uint4(0, 0, 0, 0);
        if (idx_vec && (q << 2) + 3 < n_out) return __ldg(reinterpret_cast<const uint4*>(p.iid.idx) + q);
        uint32_t iv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) iv[t] = ((q << 2) + t < n_out) ? (uint32_t)clampll(p.iid.at((q << 2) + t), p.iid_count) : 0u;
        return make_uint4(iv[0], iv[1], iv[2], iv[3]);
    };
    // ... -> word offsets and shifts (clamped: memory safe for any device index vector)
    auto decode_quad = [&](long long q, const uint4& v, uint32_t (&boff)[4], uint32_t (&sh)[4]) -> int {
        const uint32_t iv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t i = iv[t] < (uint32_t)p.iid_count ? iv[t] : (uint32_t)p.iid_count - 1u;
            boff[t] = i >> 2;
            sh[t] = 2u * (i & 3u);
        }
        return (int)min(4LL, n_out - (q << 2));
    };

    for (;;) {
        const long long batch = s_claim[iter & 1u];
        if (batch >= nbatch) break;
        const long long b0 = batch << 2;
        const int ns = (int)min(4LL, p.sid.n - b0);
        const uint8_t* src[4];
#pragma unroll
        for (int s = 0; s < 4; ++s)
            src[s] = kStaged ? (const uint8_t*)(raw0 + (size_t)s * p.raw_stride) : p.packed + clampll(p.sid.at(b0 + min(s, ns - 1)), p.sid_count) * p.ld;
        if (kStaged) mbar_wait(&stage_bar, iter & 1u);
        ++iter;
        // ---- load + interleave ----
        // per SNP: [0] selected low bits, [1] selected high bits, [2] selected fields with both bits.  Two record words share one POPC
        // (word w on the even bit positions, word w + nt on the odd ones): POPC is a 16-lane pipe and every instruction here costs an
        // issue slot of a loop that is issue-bound.
        unsigned int mc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) mc[k] = 0;
        auto load_words = [&](unsigned w, uint32_t (&x)[4]) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (kStaged) {
                    x[s] = reinterpret_cast<const uint32_t*>(src[s])[w];
                } else if (word_ok) {
                    x[s] = __ldg(reinterpret_cast<const uint32_t*>(src[s]) + w);
                } else {
                    x[s] = 0;
                    for (int k = 0; k < 4; ++k)
                        if (4 * w + k < p.rec_bytes) x[s] |= (uint32_t)__ldg(src[s] + 4 * w + k) << (8 * k);
                }
            }
        };
        auto interleave = [&](unsigned w, const uint32_t (&x)[4]) {
            const uint32_t lo01 = __byte_perm(x[0], x[1], 0x5140), lo23 = __byte_perm(x[2], x[3], 0x5140);
            const uint32_t hi01 = __byte_perm(x[0], x[1], 0x7362), hi23 = __byte_perm(x[2], x[3], 0x7362);
            reinterpret_cast<uint4*>(smem_dyn)[w] = make_uint4(__byte_perm(lo01, lo23, 0x5410), __byte_perm(lo01, lo23, 0x7632),
                                                               __byte_perm(hi01, hi23, 0x5410), __byte_perm(hi01, hi23, 0x7632));
        };
        for (unsigned w = tid; w < rec_words; w += 2 * nt) {
            const unsigned w2 = w + nt;
            const bool two = w2 < rec_words;
            uint32_t x[4], y[4] = {0u, 0u, 0u, 0u};
            load_words(w, x);
            if (two) load_words(w2, y);
            if (mask_count) {
                const uint32_t sa = __ldg(p.sel_mask + w), sb = two ? (__ldg(p.sel_mask + w2) << 1) : 0u;   // 0b01 / 0b10 per selected individual
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const uint32_t a1 = x[s] >> 1, b1 = y[s] << 1;
                    mc[3 * s] += __popc((x[s] & sa) | (b1 & sb));
                    mc[3 * s + 1] += __popc((a1 & sa) | (y[s] & sb));
                    mc[3 * s + 2] += __popc((x[s] & a1 & sa) | (y[s] & b1 & sb));
                }
            }
            interleave(w, x);
            if (two) interleave(w2, y);
        }
        if (mask_count) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {                               // code 01 = low bit only, 10 = high bit only, 11 = both
                const unsigned int n_lo = __reduce_add_sync(0xffffffffu, mc[3 * s]), n_hi = __reduce_add_sync(0xffffffffu, mc[3 * s + 1]);
                const unsigned int n_b = __reduce_add_sync(0xffffffffu, mc[3 * s + 2]);
                if ((tid & 31) == 0) {
                    if (n_lo - n_b) atomicAdd(&cnt[s][0], n_lo - n_b);
                    if (n_hi - n_b) atomicAdd(&cnt[s][1], n_hi - n_b);
                    if (n_b) atomicAdd(&cnt[s][2], n_b);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int nbt = atomicAdd(p.counter, 1);
            s_claim[iter & 1u] = nbt;                                   // iter was advanced above: the next iteration's slot, read after later barriers
            if (kStaged && nbt < nbatch) issue(nbt);                     // staging area is free again
        }
        // ---- pass 1: dosage counts of the 4 SNPs over the selected individuals ----
        if (p.mode != PSTB_STD_NONE) {
            if (!p.use_stats && !mask_count) {
                unsigned int tot[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) tot[k] = 0;
                uint32_t a1 = 0, a2 = 0, a3 = 0, pending = 0;           // 4 x 8-bit counters each (one per SNP)
                auto flush = [&]() {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        tot[3 * s] += (a1 >> (8 * s)) & 0xffu;
                        tot[3 * s + 1] += (a2 >> (8 * s)) & 0xffu;
                        tot[3 * s + 2] += (a3 >> (8 * s)) & 0xffu;
                    }
                    a1 = a2 = a3 = 0;
                    pending = 0;
                };
                uint4 nxt = fetch_quad(tid);
                for (long long q = tid; q < nq; q += nt) {
                    uint32_t boff[4], sh[4];
                    const uint4 cur = nxt;
                    nxt = fetch_quad(q + nt);
                    const int valid = decode_quad(q, cur, boff, sh);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (t < valid) {
                            const uint32_t c4 = (inter32[boff[t]] >> sh[t]) & 0x03030303u;
                            const uint32_t lo = c4 & 0x01010101u, hi = (c4 >> 1) & 0x01010101u;
                            a1 += lo & ~hi;
                            a2 += hi & ~lo;
                            a3 += hi & lo;
                        }
                    }
                    pending += 4;
                    if (pending >= 252) flush();
                }
                flush();
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const unsigned int r = __reduce_add_sync(0xffffffffu, tot[k]);
                    if ((tid & 31) == 0 && r) atomicAdd(&cnt[k / 3][k % 3], r);
                }
                __syncthreads();
            }
            if (tid < ns) {
                double mean, sd;
                if (p.use_stats) {
                    mean = p.stats[2 * (b0 + tid)];
                    sd = p.stats[2 * (b0 + tid) + 1];
                } else {
                    const long long c1 = cnt[tid][0], c2 = cnt[tid][1], c3 = cnt[tid][2], c0 = n_out - c1 - c2 - c3;
                    stats_from_counts(p.count_a1 ? c3 : c0, c2, p.count_a1 ? c0 : c3, mean, sd);
                    if (c1 && p.miss_flag) *p.miss_flag = 1u;
                    if (p.stats) {
                        p.stats[2 * (b0 + tid)] = mean;
                        p.stats[2 * (b0 + tid) + 1] = sd;
                    }
                }
                st_s[tid][0] = mean;
                st_s[tid][1] = sd;
            }
        }
        if (tid < ns) {                                                 // each thread reads only the statistics it wrote itself
            const Lut4<T> l = make_code_lut<T>(p.mode, p.count_a1, p.a, p.b, p.lnB, st_s[tid][0], st_s[tid][1]);
            lut_d[tid][0] = (double)l.c0; lut_d[tid][1] = (double)l.c1; lut_d[tid][2] = (double)l.c2; lut_d[tid][3] = (double)l.c3;
        }
        __syncthreads();
        if (tid < 12) (&cnt[0][0])[tid] = 0;                            // ready for the next batch (ordered by the barrier below)
        // ---- pass 2: emit ----
        if (p.out) {
            Lut4<T> lut[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) { lut[s].c0 = (T)lut_d[s][0]; lut[s].c1 = (T)lut_d[s][1]; lut[s].c2 = (T)lut_d[s][2]; lut[s].c3 = (T)lut_d[s][3]; }
            T* o = reinterpret_cast<T*>(p.out) + b0 * p.out_ld;
            uint4 nxt = fetch_quad(tid);
            for (long long q = tid; q < nq; q += nt) {
                uint32_t boff[4], sh[4], c4[4];
                const uint4 cur = nxt;
                nxt = fetch_quad(q + nt);
                const int valid = decode_quad(q, cur, boff, sh);
#pragma unroll
                for (int t = 0; t < 4; ++t) c4[t] = (inter32[boff[t]] >> sh[t]) & 0x03030303u;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (s < ns) {
                        T* oc = o + (long long)s * p.out_ld + (q << 2);
                        const T v0 = lut[s].pick((c4[0] >> (8 * s)) & 3u), v1 = lut[s].pick((c4[1] >> (8 * s)) & 3u);
                        const T v2 = lut[s].pick((c4[2] >> (8 * s)) & 3u), v3 = lut[s].pick((c4[3] >> (8 * s)) & 3u);
                        if (valid == 4) store4<T>(oc, v0, v1, v2, v3, p.vec_ok);
                        else {
                            if (valid > 0) oc[0] = v0;
                            if (valid > 1) oc[1] = v1;
                            if (valid > 2) oc[2] = v2;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---- C order: tile-transposing emit ------------------------------------------------------------------
// A tile is (32 * V) SNPs x 512 individuals; every lane owns V adjacent SNPs, so a warp writes one row segment of
// 32 * V * sizeof(T) contiguous bytes per store (256 B for float32 with V = 2 and for float64 with V = 1).
constexpr int kTileS = 32;    // lanes across SNPs
constexpr int kTileI = 512;   // individuals per tile

template <typename T, int V>
struct VecStore;
template <typename T>
struct VecStore<T, 1> {
    static __device__ __forceinline__ void st(T* o, const T* v) { __stcs(o, v[0]); }
};
template <>
struct VecStore<float, 2> {
    static __device__ __forceinline__ void st(float* o, const float* v) { __stcs(reinterpret_cast<float2*>(o), make_float2(v[0], v[1])); }
};
template <>
struct VecStore<int8_t, 4> {
    static __device__ __forceinline__ void st(int8_t* o, const int8_t* v) {
        *reinterpret_cast<uint32_t*>(o) = (uint32_t)(uint8_t)v[0] | ((uint32_t)(uint8_t)v[1] << 8) | ((uint32_t)(uint8_t)v[2] << 16) | ((uint32_t)(uint8_t)v[3] << 24);
    }
};

template <typename T, int V>
__global__ void __launch_bounds__(256) k_emit_c(const ReadParams p, long long tiles_i) {
    constexpr int TS = kTileS * V;
    constexpr int kPitch = kTileI / 4 + 4;          // 33 words: word-aligned rows; lanes V apart cost at most a V-way bank conflict
    __shared__ __align__(16) unsigned char codes[TS][kPitch];
    __shared__ T lut_s[TS][4];
    // consecutive CTAs take consecutive 512-individual blocks of the SAME records: their 128-byte packed reads are adjacent
    // (the SNP-block-fastest order was measured slower: 43.7 % vs 47.3 % of the HBM peak for float32)
    const long long tile = blockIdx.x;
    const long long ts = tile / tiles_i, ti = tile % tiles_i;
    const long long b0 = ts * TS, i0 = ti * kTileI;
    const long long n_out = p.iid.n;
    const int rows = (int)min((long long)kTileI, n_out - i0);
    const int nsnp = (int)min((long long)TS, p.sid.n - b0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    // packed bytes of the tile -> shared memory (one byte = 4 consecutive output individuals); one warp per record:
    // 32 lanes x 32-bit loads = the record's 128-byte fragment in one coalesced request
    const int nbytes = (rows + 3) >> 2;
    const bool word_ok = p.dense && (p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.packed) & 3u) == 0) && (p.byte_off % 4 == 0);
    for (int s = warp; s < nsnp; s += nwarps) {
        const long long j = clampll(p.sid.at(b0 + s), p.sid_count);
        const uint8_t* src = p.packed + j * p.ld;
        if (word_ok) {
            // the fragment may run up to 3 bytes past the record inside its ld padding; those codes are never emitted
            const uint8_t* frag = src + p.byte_off + (i0 >> 2);
            if (4 * lane < nbytes) *reinterpret_cast<uint32_t*>(&codes[s][4 * lane]) = __ldg(reinterpret_cast<const uint32_t*>(frag) + lane);
        } else {
            for (int q = lane; q < nbytes; q += 32) {
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
        }
    }
    for (int s = threadIdx.x; s < nsnp; s += blockDim.x) {
        double mean = 0.0, sd = 1.0;
        if (p.mode != PSTB_STD_NONE) {
            mean = p.stats[2 * (b0 + s)];
            sd = p.stats[2 * (b0 + s) + 1];
        }
        Lut4<T> l = make_code_lut<T>(p.mode, p.count_a1, p.a, p.b, p.lnB, mean, sd);
        lut_s[s][0] = l.c0; lut_s[s][1] = l.c1; lut_s[s][2] = l.c2; lut_s[s][3] = l.c3;
    }
    __syncthreads();
    if (V * lane >= nsnp) return;
    Lut4<T> lut[V];
    const unsigned char* crow[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int s = min(V * lane + k, nsnp - 1);
        lut[k].c0 = lut_s[s][0]; lut[k].c1 = lut_s[s][1]; lut[k].c2 = lut_s[s][2]; lut[k].c3 = lut_s[s][3];
        crow[k] = codes[s];
    }
    const bool full = V * lane + V <= nsnp;
    // one shared-memory byte serves 4 consecutive rows: each warp takes row quads q = warp, warp + nwarps, ...
    T* dst = reinterpret_cast<T*>(p.out) + (i0 + 4LL * warp) * p.out_ld + b0 + V * lane;
    const long long dstep = 4LL * nwarps * p.out_ld;
    for (int q = warp; q < nbytes; q += nwarps, dst += dstep) {
        uint32_t byte[V];
#pragma unroll
        for (int k = 0; k < V; ++k) byte[k] = crow[k][q];
        const int nr = min(4, rows - 4 * q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (t < nr) {
                T v[V];
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] = lut[k].pick((byte[k] >> (2 * t)) & 3u);
                T* d = dst + (long long)t * p.out_ld;
                if (full) VecStore<T, V>::st(d, v);
                else
                    for (int k = 0; k < V; ++k)
                        if (V * lane + k < nsnp) d[k] = v[k];
            }
        }
    }
}

// ---- C order, wide rows: 32 bytes (one 256-bit store) per lane, 1 KiB of one output row per warp store ------------------
// The tile is (32 * V) SNPs x 512 individuals with V = 32 / sizeof(T) adjacent SNPs per lane (float32: 256 SNPs, float64: 128).
// DRAM sees 1 KiB contiguous row pieces instead of 256 B ones.  The packed bytes are transposed on the way into shared
// memory -- codes_t[row quad][SNP] -- so a lane fetches the codes of its V SNPs for four rows with ONE 64- / 32-bit
// shared load, conflict-free (consecutive lanes read consecutive bytes).  Lane <-> record during the fill: the 32 lanes of a
// warp walk the 128-byte fragments of 32 records word by word (each fragment is one L1 line, touched 32 times) and drop
// the four bytes of a word into four smem rows, 32 consecutive bytes per warp store.
template <typename T>
struct WideRow;
template <>
struct WideRow<float> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void st(float* o, const float (&v)[8]) { st256_f32(o, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]); }
};
template <>
struct WideRow<double> {
    static constexpr int V = 4;
    static __device__ __forceinline__ void st(double* o, const double (&v)[4]) { st256_f64(o, v[0], v[1], v[2], v[3]); }
};

template <typename T>
__global__ void __launch_bounds__(256) k_emit_c_wide(const ReadParams p, long long tiles_i) {
    constexpr int V = WideRow<T>::V;
    constexpr int TS = kTileS * V;                  // SNPs per tile
    constexpr int Q = kTileI / 4;                   // row quads per tile
    constexpr int kPitch = TS + 8;                  // bytes per row quad (8-byte aligned rows for the 64-bit loads)
    __shared__ __align__(16) unsigned char codes_t[Q][kPitch];
    __shared__ T lut_s[TS][4];
    const long long tile = blockIdx.x;
    const long long ts = tile / tiles_i, ti = tile % tiles_i;
    const long long b0 = ts * TS, i0 = ti * kTileI;
    const long long n_out = p.iid.n;
    const int rows = (int)min((long long)kTileI, n_out - i0);
    const int nsnp = (int)min((long long)TS, p.sid.n - b0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nbytes = (rows + 3) >> 2;             // row quads in use
    const bool word_ok = p.dense && (p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.packed) & 3u) == 0) && (p.byte_off % 4 == 0);

    for (int s0 = warp * 32; s0 < TS; s0 += nwarps * 32) {
        const int s = s0 + lane;
        const bool live = s < nsnp;
        const long long j = live ? clampll(p.sid.at(b0 + s), p.sid_count) : 0;
        const uint8_t* src = p.packed + j * p.ld;
        if (word_ok && nbytes == Q && p.ld % 16 == 0 && ((reinterpret_cast<uintptr_t>(p.packed) + (uintptr_t)p.byte_off) & 15u) == 0) {
            // full tile: the whole 128-byte fragment with eight independent 128-bit loads (all in flight at once: the fill was
            // the exposed latency of this kernel, stall long_scoreboard 15 per issue in profiles/r1_late_kernels_full.txt)
            const uint4* frag = reinterpret_cast<const uint4*>(src + p.byte_off + (i0 >> 2));
            uint4 v[Q / 16];
#pragma unroll
            for (int k = 0; k < Q / 16; ++k) v[k] = live ? __ldg(frag + k) : make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u);
#pragma unroll
            for (int k = 0; k < Q / 16; ++k) {
                const uint32_t wd[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int t = 0; t < 4; ++t) codes_t[16 * k + 4 * u + t][s] = (unsigned char)(wd[u] >> (8 * t));
            }
        } else if (word_ok) {
            const uint32_t* frag = reinterpret_cast<const uint32_t*>(src + p.byte_off + (i0 >> 2));
            const int nwords = (nbytes + 3) >> 2;   // may run up to 3 bytes into the record's ld padding; those codes are never emitted
#pragma unroll 4
            for (int w = 0; w < nwords; ++w) {
                const uint32_t word = live ? __ldg(frag + w) : 0x55555555u;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (4 * w + t < Q) codes_t[4 * w + t][s] = (unsigned char)(word >> (8 * t));
            }
        } else {
            for (int q = 0; q < nbytes; ++q) {
                uint32_t byte = 0x55u;
                if (live) {
                    if (p.dense) {
                        byte = __ldg(src + p.byte_off + (i0 >> 2) + q);
                    } else {
                        byte = 0;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const long long a = i0 + 4 * q + t;
                            if (a < n_out) {
                                const long long i = clampll(p.iid.at(a), p.iid_count);
                                byte |= ((uint32_t)(__ldg(src + (i >> 2)) >> (2 * (i & 3))) & 3u) << (2 * t);
                            }
                        }
                    }
                }
                codes_t[q][s] = (unsigned char)byte;
            }
        }
    }
    for (int s = threadIdx.x; s < nsnp; s += blockDim.x) {
        double mean = 0.0, sd = 1.0;
        if (p.mode != PSTB_STD_NONE) {
            mean = p.stats[2 * (b0 + s)];
            sd = p.stats[2 * (b0 + s) + 1];
        }
        Lut4<T> l = make_code_lut<T>(p.mode, p.count_a1, p.a, p.b, p.lnB, mean, sd);
        lut_s[s][0] = l.c0; lut_s[s][1] = l.c1; lut_s[s][2] = l.c2; lut_s[s][3] = l.c3;
    }
    __syncthreads();
    if (V * lane >= nsnp) return;
    Lut4<T> lut[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int s = min(V * lane + k, nsnp - 1);
        lut[k].c0 = lut_s[s][0]; lut[k].c1 = lut_s[s][1]; lut[k].c2 = lut_s[s][2]; lut[k].c3 = lut_s[s][3];
    }
    const bool full = V * lane + V <= nsnp;
    T* dst = reinterpret_cast<T*>(p.out) + (i0 + 4LL * warp) * p.out_ld + b0 + V * lane;
    const long long dstep = 4LL * nwarps * p.out_ld;
    for (int q = warp; q < nbytes; q += nwarps, dst += dstep) {
        uint32_t lo, hi = 0;
        if (V == 8) {
            const uint2 two = *reinterpret_cast<const uint2*>(&codes_t[q][V * lane]);
            lo = two.x;
            hi = two.y;
        } else {
            lo = *reinterpret_cast<const uint32_t*>(&codes_t[q][V * lane]);
        }
        const int nr = min(4, rows - 4 * q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (t < nr) {
                T v[V];
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const uint32_t byte = ((k < 4 ? lo : hi) >> (8 * (k & 3))) & 0xffu;
                    v[k] = lut[k].pick((byte >> (2 * t)) & 3u);
                }
                T* d = dst + (long long)t * p.out_ld;
                if (full) WideRow<T>::st(d, v);
                else
                    for (int k = 0; k < V; ++k)
                        if (V * lane + k < nsnp) d[k] = v[k];
            }
        }
    }
}

// returns -1 when the wide kernel does not apply (int8, rows not 32-byte aligned, fewer SNPs than one tile): narrow tiles then
template <typename T>
static int launch_emit_c_wide(const ReadParams& p, cudaStream_t st) {
    if constexpr (sizeof(T) == 1) {
        return -1;
    } else {
        constexpr int V = WideRow<T>::V, TS = kTileS * V;
        static const bool narrow_env = getenv("PSTB_EMIT_C_NARROW") && atoi(getenv("PSTB_EMIT_C_NARROW")) != 0;     // A/B runs
        if (narrow_env || p.sid.n < TS || p.sid.n % V != 0 || (reinterpret_cast<uintptr_t>(p.out) & 31u)) return -1;
        const long long tiles_i = (p.iid.n + kTileI - 1) / kTileI, tiles_s = (p.sid.n + TS - 1) / TS;
        const long long tiles = tiles_i * tiles_s;
        if (tiles > 0x7fffffffLL) return fail("C-order read too large for one launch (%lld tiles)", tiles);
        k_emit_c_wide<T><<<(unsigned)tiles, 256, 0, st>>>(p, tiles_i);
        PSTB_AFTER_LAUNCH("k_emit_c_wide");
        return 0;
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
        {
            const int rc = launch_emit_c_wide<T>(p, st);
            if (rc >= 0) return rc;
        }
        // V adjacent SNPs per lane when the rows allow vector stores (row stride and base aligned to V elements)
        constexpr int kV = sizeof(T) == 4 ? 2 : (sizeof(T) == 1 ? 4 : 1);
        const bool vec = kV > 1 && (p.sid.n % kV == 0) && (reinterpret_cast<uintptr_t>(p.out) % (kV * sizeof(T)) == 0);
        const int V = vec ? kV : 1;
        const long long tiles_i = (n_out + kTileI - 1) / kTileI, tiles_s = (p.sid.n + kTileS * V - 1) / (kTileS * V);
        const long long tiles = tiles_i * tiles_s;
        if (tiles > 0x7fffffffLL) return fail("C-order read too large for one launch (%lld tiles)", tiles);
        if (vec) k_emit_c<T, kV><<<(unsigned)tiles, 256, 0, st>>>(p, tiles_i);
        else k_emit_c<T, 1><<<(unsigned)tiles, 256, 0, st>>>(p, tiles_i);
        PSTB_AFTER_LAUNCH("k_emit_c");
        return 0;
    }
    p.out_ld = n_out;
    p.vec_ok = 0;
    if (next_counter(st, &p.counter)) return 1;                   // every F-order kernel below hands its records out through it
    if (p.out) {
        const uintptr_t base = reinterpret_cast<uintptr_t>(p.out);
        const long long col = n_out * (long long)sizeof(T);
        p.vec_ok = ((base & 31u) == 0 && col % 32 == 0) ? 2 : (((base & 15u) == 0 && col % 16 == 0) ? 1 : 0);
    }
    const unsigned rec16 = (p.rec_bytes + 15u) & ~15u;
    if (!p.out && p.stats && (reinterpret_cast<uintptr_t>(p.stats) & 15u) == 0 && p.mode != PSTB_STD_NONE && !p.use_stats && p.dense &&
        n_out >= 1024 && p.ld % 16 == 0 &&
        ((reinterpret_cast<uintptr_t>(p.packed) + (uintptr_t)p.byte_off) & 15u) == 0 &&
        ((n_out + 15) / 16 * 4 + p.byte_off <= p.ld || p.sid_count == 0) && !getenv("PSTB_STATS_V1")) {
        // statistics only: 128-bit register streaming (the tail's last 32-bit word stays inside the record's ld bytes)
        long long grid = (long long)sms * 8;
        int R = 1;                                               // records per warp block: ~2 blocks per resident warp, at most 32
        while (R < 32 && p.sid.n >= grid * 8 * 2 * (2 * R)) R *= 2;
        const long long want = (p.sid.n + 8LL * R - 1) / (8LL * R);
        if (grid > want) grid = want;
        if (grid < 1) grid = 1;
        k_stats_dense<<<(unsigned)grid, 256, 0, st>>>(p, R);
        PSTB_AFTER_LAUNCH("k_stats_dense");
        return 0;
    }
    p.bulk_ok = ((reinterpret_cast<uintptr_t>(p.packed) & 15u) == 0) && (p.ld % 16 == 0) && (rec16 <= p.ld || p.sid_count == 0);
    p.copy_bytes = rec16;
    p.raw_stride = rec16;
    const unsigned dense_bytes = p.dense ? 0u : (unsigned)((((n_out + 3) >> 2) + 15) & ~15LL);
    const unsigned max_smem = 220u * 1024u;
    if (!p.dense) {
        const unsigned inter_bytes = 16u * ((p.rec_bytes + 3u) >> 2);
        if (inter_bytes <= max_smem && !getenv("PSTB_GATHER_V1")) {
            const long long nq = (n_out + 3) >> 2;
            const long long nbatch = (p.sid.n + 3) / 4;
            const unsigned staged_bytes = inter_bytes + 4u * rec16;
            const bool staged = p.bulk_ok && staged_bytes <= max_smem && !getenv("PSTB_GATHER_NOSTAGE");
            if (staged) {
                keep_pool_memory(st);
                uint32_t* d_mask = nullptr;
                if (p.mode != PSTB_STD_NONE && !p.use_stats) {
                    p.mask_words = (p.iid_count + 15) / 16;
                    if (cudaMallocAsync(reinterpret_cast<void**>(&d_mask), (size_t)(p.mask_words + 1) * 4, st) == cudaSuccess) {
                        PSTB_CUDA(cudaMemsetAsync(d_mask, 0, (size_t)(p.mask_words + 1) * 4, st));
                        long long g = (n_out + 255) / 256;
                        if (g > (long long)sms * 8) g = (long long)sms * 8;
                        k_build_sel_mask<<<(unsigned)g, 256, 0, st>>>(p.iid, p.iid_count, d_mask, p.mask_words);
                        PSTB_AFTER_LAUNCH("k_build_sel_mask");
                        p.sel_mask = d_mask;
                    } else {
                        cudaGetLastError();                       // no pool memory: count through the gather instead
                    }
                }
                // one CTA per SM: TMA-staged raw records for batch b+1 while batch b is counted and written
                int threads = nq >= 1024 ? 1024 : (int)(((nq + 31) / 32) * 32);
                if (threads < 64) threads = 64;
                PSTB_CUDA(cudaFuncSetAttribute(k_read_f_gather4<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_bytes));
                int ctas_per_sm = 1;
                PSTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_read_f_gather4<T, true>, threads, staged_bytes));
                if (ctas_per_sm < 1) ctas_per_sm = 1;
                long long grid = (long long)sms * ctas_per_sm;
                if (grid > nbatch) grid = nbatch;
                k_read_f_gather4<T, true><<<(unsigned)grid, threads, staged_bytes, st>>>(p);
                PSTB_AFTER_LAUNCH("k_read_f_gather4<staged>");
                if (d_mask) PSTB_CUDA(cudaFreeAsync(d_mask, st));
                return 0;
            }
            // 4 byte-interleaved records per batch loaded straight from global; two CTAs per SM when the tile is <= ~110 KiB
            int threads = nq >= 512 ? 512 : (int)(((nq + 31) / 32) * 32);
            if (threads < 64) threads = 64;
            PSTB_CUDA(cudaFuncSetAttribute(k_read_f_gather4<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inter_bytes));
            int ctas_per_sm = 1;
            PSTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_read_f_gather4<T, false>, threads, inter_bytes));
            if (ctas_per_sm < 1) ctas_per_sm = 1;
            long long grid = (long long)sms * ctas_per_sm;
            if (grid > nbatch) grid = nbatch;
            k_read_f_gather4<T, false><<<(unsigned)grid, threads, inter_bytes, st>>>(p);
            PSTB_AFTER_LAUNCH("k_read_f_gather4");
            return 0;
        }
        // larger records: batch S records per pass over the index vector when at least one fits
        const unsigned per_snp = rec16 + dense_bytes;
        int S = (int)((max_smem - 16u) / per_snp);
        if (S > kGatherMaxS) S = kGatherMaxS;
        if (S >= 1) {
            // two CTAs per SM (the re-pack of one overlaps the HBM writes of the other) beat one big CTA: measured on
            // cfg4 S=3 x 512 threads x 2 CTAs = 67 % of the HBM roofline vs 58 % for S=4 x 1024 threads x 1 CTA
            int threads = n_out < 16384 ? 256 : 512;
            const unsigned half_sm = (227u * 1024u) / 2u - 1024u - 16u;
            while (S > 1 && (unsigned)S * per_snp > half_sm) --S;
            while (S > 4 && (unsigned)S * per_snp > 64u * 1024u) --S;      // small records: room for several CTAs per SM
            if (const char* e = getenv("PSTB_GATHER_S")) { int v = atoi(e); if (v >= 1 && v <= S) S = v; }          // tuning experiments
            if (const char* e = getenv("PSTB_GATHER_THREADS")) { int v = atoi(e); if (v >= 64 && v <= 1024 && v % 32 == 0) threads = v; }
            if ((long long)S > p.sid.n) S = (int)p.sid.n;
            const unsigned smem = 16u + (unsigned)S * per_snp;
            PSTB_CUDA(cudaFuncSetAttribute(k_read_f_gather<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int ctas_per_sm = 1;
            PSTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_read_f_gather<T>, threads, smem));
            if (ctas_per_sm < 1) ctas_per_sm = 1;
            const long long nbatch = (p.sid.n + S - 1) / S;
            long long grid = (long long)sms * ctas_per_sm;
            if (grid > nbatch) grid = nbatch;
            k_read_f_gather<T><<<(unsigned)grid, threads, smem, st>>>(p, S, dense_bytes);
            PSTB_AFTER_LAUNCH("k_read_f_gather");
            return 0;
        }
    }
    // short records, adjacent in the store: batches of R records per warp
    if (p.dense && p.bulk_ok && p.ld <= kSmallLd && p.sid.idx == nullptr && p.sid.step == 1 && p.byte_off % 4 == 0 &&
        (p.mode == PSTB_STD_NONE || p.use_stats || p.byte_off + 4 * ((n_out + 15) / 16) <= p.ld) && !getenv("PSTB_READ_SMALL_V1")) {
        int R = 32;
        while (R > 1 && (long long)R * p.ld > 4096) R >>= 1;
        const int warps = 8;
        const unsigned smem = warps * (16u + 2u * (unsigned)R * (unsigned)p.ld);
        PSTB_CUDA(cudaFuncSetAttribute(k_read_f_small<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int ctas_per_sm = 1;
        PSTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_read_f_small<T>, warps * 32, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        const long long nbatch = (p.sid.n + R - 1) / R;
        long long grid = (long long)sms * ctas_per_sm;
        if (grid > (nbatch + warps - 1) / warps) grid = (nbatch + warps - 1) / warps;
        if (grid < 1) grid = 1;
        k_read_f_small<T><<<(unsigned)grid, warps * 32, smem, st>>>(p, R);
        PSTB_AFTER_LAUNCH("k_read_f_small");
        return 0;
    }
    // warp-per-record when two raw buffers + the gather record stay small, else CTA-per-record
    const unsigned warp_group = 16u + 2u * rec16 + dense_bytes;
    if (warp_group <= 12u * 1024u) {
        p.nbuf = 2;
        p.group_smem = warp_group;
        // eight resident warps per SM write DRAM best; columns up to 20 KB (int8 output, N < ~5 000 float32) want them as two CTAs of four
        // (cfg2 int8: 4 x 2 89.7 % of the copy peak, 8 x 1 80.9 %, 8 x 2 and more 60 %; float32 N = 4 104: 93.2 vs 90.9 %)
        const bool long_out = p.out && p.rec_bytes >= 1024;
        const bool two_ctas = long_out && n_out * (long long)sizeof(T) <= 20 * 1024;
        int warps = two_ctas ? 4 : 8;
        if (const char* e = getenv("PSTB_READ_WARPS")) { int v = atoi(e); if (v >= 1 && v <= 8) warps = v; }      // tuning experiments
        const unsigned smem = warps * p.group_smem;
        PSTB_CUDA(cudaFuncSetAttribute(k_read_f<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // persistent grid of resident CTAs; the warps claim their records from p.counter.  Eight resident warps per SM write DRAM
        // best (cfg2, scripts/bench_read.py feed: 8 warps x 1 CTA 101.8 % of the measured copy peak, 6 x 1 102.8 %, 4 x 2 101.5 %, 8 x 2
        // 98.1 %, 4 x 4 98.0 %; 4 x 1 is too few: 76.8 %).  Short records keep full occupancy.  A statistics-only pass (no output) is a
        // pure latency-bound read and wants every resident warp it can get.
        int ctas_per_sm = 1;
        PSTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_read_f<T, false>, warps * 32, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        if (long_out) ctas_per_sm = two_ctas && ctas_per_sm >= 2 ? 2 : 1;
        if (const char* e = getenv("PSTB_READ_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 8) ctas_per_sm = v; }   // tuning experiments
        long long want = (p.sid.n + warps - 1) / warps;
        long long grid = (long long)sms * ctas_per_sm;
        if (grid > want) grid = want;
        if (grid < 1) grid = 1;
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
        PSTB_CUDA(cudaFuncSetAttribute(k_read_f<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.group_smem));
        int ctas_per_sm = 1;
        PSTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_read_f<T, true>, 512, p.group_smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        long long grid = (long long)sms * ctas_per_sm;
        if (grid > p.sid.n) grid = p.sid.n;
        if (grid < 1) grid = 1;
        k_read_f<T, true><<<(unsigned)grid, 512, p.group_smem, st>>>(p);
        PSTB_AFTER_LAUNCH("k_read_f<cta>");
    }
    return 0;
}

int read_impl_ex(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                 int count_a1, int mode, double a, double b, int use_stats, double* d_stats, void* d_out, int dtype, int order,
                 void* stream, unsigned int* d_miss_flag);

int read_impl(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
              int count_a1, int mode, double a, double b, int use_stats, double* d_stats, void* d_out, int dtype, int order,
              void* stream) {
    return read_impl_ex(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, use_stats, d_stats, d_out, dtype, order, stream,
                        nullptr);
}

int read_impl_ex(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                 int count_a1, int mode, double a, double b, int use_stats, double* d_stats, void* d_out, int dtype, int order,
                 void* stream, unsigned int* d_miss_flag) {
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
    p.miss_flag = d_miss_flag;
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

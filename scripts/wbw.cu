// write-bandwidth pattern probe (experiment; not part of the product)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
__global__ void fill_linear(float4* o, size_t n4) {
  for (size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; i < n4; i += (size_t)gridDim.x*blockDim.x) __stcs(o+i, make_float4(1,2,3,4));
}
// each warp owns column b (col4 float4s), grid-stride over columns
__global__ void fill_warpcol(float4* o, size_t ncol, int col4) {
  int lane = threadIdx.x & 31; size_t w = ((size_t)blockIdx.x*blockDim.x+threadIdx.x)>>5, nw = ((size_t)gridDim.x*blockDim.x)>>5;
  for (size_t b = w; b < ncol; b += nw) { float4* c = o + b*col4;
#pragma unroll 4
    for (int q = lane; q < col4; q += 32) __stcs(c+q, make_float4(1,2,3,(float)q)); }
}
// each CTA owns column b
__global__ void fill_ctacol(float4* o, size_t ncol, int col4) {
  for (size_t b = blockIdx.x; b < ncol; b += gridDim.x) { float4* c = o + b*col4;
#pragma unroll 4
    for (int q = threadIdx.x; q < col4; q += blockDim.x) __stcs(c+q, make_float4(1,2,3,(float)q)); }
}
// warp columns + a read of col4/16 float4 (6%) per column from a second buffer, with a dependency like ours
__global__ void fill_warpcol_rd(float4* o, const uint4* in, size_t ncol, int col4) {
  int lane = threadIdx.x & 31; size_t w = ((size_t)blockIdx.x*blockDim.x+threadIdx.x)>>5, nw = ((size_t)gridDim.x*blockDim.x)>>5;
  int in4 = col4/16;
  for (size_t b = w; b < ncol; b += nw) { float4* c = o + b*col4; const uint4* r = in + b*in4; unsigned acc=0;
    for (int q = lane; q < in4; q += 32) { uint4 v = __ldg(r+q); acc += __popc(v.x)+__popc(v.y)+__popc(v.z)+__popc(v.w); }
    for (int s=16;s;s>>=1) acc += __shfl_xor_sync(~0u, acc, s);
    float f = (float)acc;
#pragma unroll 4
    for (int q = lane; q < col4; q += 32) __stcs(c+q, make_float4(f,2,3,(float)q)); }
}
// clumped reads: the CTA loads the packed records of NB adjacent columns in one go (contiguous NB*in4 uint4), syncs,
// then each warp writes its columns.  Same bytes as fill_warpcol_rd, different read granularity.
template <int NB>
__global__ void fill_ctabatch_rd(float4* o, const uint4* in, size_t ncol, int col4) {
  extern __shared__ uint4 sm[];
  int in4 = col4/16; int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (size_t b0 = (size_t)blockIdx.x * NB; b0 < ncol; b0 += (size_t)gridDim.x * NB) {
    const uint4* r = in + b0*in4;
    for (int q = threadIdx.x; q < NB*in4; q += blockDim.x) sm[q] = __ldg(r+q);
    __syncthreads();
    for (int c = warp; c < NB; c += nw) { unsigned acc=0;
      for (int q = lane; q < in4; q += 32) { uint4 v = sm[c*in4+q]; acc += __popc(v.x)+__popc(v.y)+__popc(v.z)+__popc(v.w); }
      for (int s=16;s;s>>=1) acc += __shfl_xor_sync(~0u, acc, s);
      float f = (float)acc; float4* cc = o + (b0+c)*col4;
#pragma unroll 4
      for (int q = lane; q < col4; q += 32) __stcs(cc+q, make_float4(f,2,3,(float)q)); }
    __syncthreads();
  }
}
// V_a: our emit loop (LDS.U16 + 8 register selects + 256-bit store) on a static smem record: no global reads at all
__device__ __forceinline__ float pick4(unsigned c, float a, float b, float cc, float d) { float lo = (c & 1u) ? b : a; float hi = (c & 1u) ? d : cc; return (c & 2u) ? hi : lo; }
__global__ void emit_like(float* o, size_t ncol, int col_f, int prefetch, const unsigned char* in, int rec16) {
  extern __shared__ __align__(16) unsigned char smx[];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5; size_t gw = ((size_t)blockIdx.x*blockDim.x+threadIdx.x)>>5, nw = ((size_t)gridDim.x*blockDim.x)>>5;
  unsigned char* raw = smx + (size_t)w * (2*rec16 + 16);
  unsigned long long* bar = (unsigned long long*)(raw + 2*rec16);
  unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar);
  for (int i = lane; i < 2*rec16; i += 32) raw[i] = (unsigned char)(i * 37 + w);
  if (prefetch && lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncwarp();
  unsigned it = 0;
  if (prefetch && lane == 0 && gw < ncol) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(rec16)); asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"((unsigned)__cvta_generic_to_shared(raw)), "l"(in + gw*rec16), "r"(rec16), "r"(bar_a) : "memory"); }
  const int WIN = 16;   // iterations per L2 prefetch window
  for (size_t b = gw; b < ncol; b += nw, ++it) {
    unsigned char* rec = raw + (it & 1) * rec16;
    if (prefetch == 2 && gw == 0 && (it % WIN) == 0) {
      // burst: pull the packed records of iterations [it+WIN, it+2*WIN) into L2 in one go (first window: also [it, it+WIN))
      size_t lo = (size_t)(it + (it == 0 ? 0 : WIN)) * nw * rec16, hi = (size_t)(it + 2 * WIN) * nw * rec16, tot = ncol * (size_t)rec16;
      if (hi > tot) hi = tot;
      for (size_t off = lo + (size_t)lane * 32768; off < hi; off += 32 * 32768) {
        unsigned sz = (unsigned)((hi - off) < 32768 ? (hi - off) : 32768); sz &= ~15u;
        if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(in + off), "r"(sz) : "memory");
      }
    }
    if (prefetch == 3 && (it % WIN) == 0) {
      // every warp pulls its own next WIN records into L2 at the same moment: a chip-wide read burst every WIN iterations
      for (int k = lane; k < (it == 0 ? 2 * WIN : WIN); k += 32) {
        size_t bb = b + (size_t)(k + (it == 0 ? 0 : WIN)) * nw;
        if (bb < ncol) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(in + bb * rec16), "r"(rec16) : "memory");
      }
    }
    if (prefetch) {
      unsigned ok = 0; while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(bar_a), "r"(it & 1) : "memory");
      if (lane == 0 && b + nw < ncol) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(rec16)); asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"((unsigned)__cvta_generic_to_shared(raw + ((it+1)&1)*rec16)), "l"(in + (b+nw)*rec16), "r"(rec16), "r"(bar_a) : "memory"); }
    }
    float l0 = (float)b, l1 = 0.f, l2 = l0 + 1.f, l3 = l0 + 2.f;
    const unsigned short* r16 = (const unsigned short*)rec; float* c = o + b * (size_t)col_f; int nh = col_f >> 3;
#pragma unroll 4
    for (int h = lane; h < nh; h += 32) { unsigned two = r16[h];
      float a0=pick4(two&3,l0,l1,l2,l3),a1=pick4((two>>2)&3,l0,l1,l2,l3),a2=pick4((two>>4)&3,l0,l1,l2,l3),a3=pick4((two>>6)&3,l0,l1,l2,l3),a4=pick4((two>>8)&3,l0,l1,l2,l3),a5=pick4((two>>10)&3,l0,l1,l2,l3),a6=pick4((two>>12)&3,l0,l1,l2,l3),a7=pick4(two>>14,l0,l1,l2,l3);
      asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(c + 8*h), "f"(a0),"f"(a1),"f"(a2),"f"(a3),"f"(a4),"f"(a5),"f"(a6),"f"(a7) : "memory"); }
    __syncwarp();
  }
}
template<class F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); f(); cudaDeviceSynchronize(); float best=1e9; for(int i=0;i<5;i++){cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} return best; }
int main(){
  size_t ncol = 1000000; int col4 = 2500; size_t n4 = ncol*col4; float4* o; uint4* in;
  CK(cudaMalloc(&o, n4*16)); CK(cudaMalloc(&in, ncol*(col4/16)*16)); CK(cudaMemset(in, 0x5a, ncol*(col4/16)*16));
  double gb = n4*16/1e9;
  for (int bps : {2,4,8,16}) { float ms = timeit([&]{ fill_linear<<<148*bps,256>>>(o,n4); }); printf("linear grid=148x%d x256: %.3f ms %.0f GB/s\n", bps, ms, gb/ms*1e3); }
  for (int bps : {1,2,4,5,8}) { float ms = timeit([&]{ fill_warpcol<<<148*bps,256>>>(o,ncol,col4); }); printf("warp-column 40KB grid=148x%d x256: %.3f ms %.0f GB/s\n", bps, ms, gb/ms*1e3); }
  for (int bps : {1,2,4,8}) { float ms = timeit([&]{ fill_ctacol<<<148*bps,256>>>(o,ncol,col4); }); printf("cta-column 40KB grid=148x%d x256: %.3f ms %.0f GB/s\n", bps, ms, gb/ms*1e3); }
  for (int bps : {2,4,8}) { float ms = timeit([&]{ fill_warpcol_rd<<<148*bps,256>>>(o,in,ncol,col4); }); printf("warp-column + 6%% dependent read grid=148x%d: %.3f ms %.0f GB/s (w+r)\n", bps, ms, gb*1.0625/ms*1e3); }
  { auto k = fill_ctabatch_rd<32>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32*(col4/16)*16);
    for (int thr : {256, 512, 1024}) for (int bps : {1,2}) { float ms = timeit([&]{ k<<<148*bps,thr,32*(col4/16)*16>>>(o,in,ncol,col4); }); printf("cta-batch32 clumped read, grid=148x%d x%d: %.3f ms %.0f GB/s (w+r)\n", bps, thr, ms, gb*1.0625/ms*1e3); } }
  { auto k = fill_ctabatch_rd<8>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8*(col4/16)*16);
    for (int bps : {2,4,8}) { float ms = timeit([&]{ k<<<148*bps,256,8*(col4/16)*16>>>(o,in,ncol,col4); }); printf("cta-batch8 clumped read, grid=148x%d x256: %.3f ms %.0f GB/s (w+r)\n", bps, ms, gb*1.0625/ms*1e3); } }
  { int rec16 = 2512; unsigned char* pk; cudaMalloc(&pk, ncol*(size_t)rec16); cudaMemset(pk, 0x9c, ncol*(size_t)rec16);
    cudaFuncSetAttribute(emit_like, cudaFuncAttributeMaxDynamicSharedMemorySize, 8*(2*rec16+16));
    for (int pf : {0, 1, 3}) for (int bps : {1,2,4}) { float ms = timeit([&]{ emit_like<<<148*bps,256,8*(2*rec16+16)>>>((float*)o,ncol,10000,pf,pk,rec16); }); printf("emit-like (LDS+select+st.v8) prefetch=%d grid=148x%d: %.3f ms %.0f GB/s\n", pf, bps, ms, (gb + (pf? ncol*2512e-9:0))/ms*1e3); } }
  return 0;
}

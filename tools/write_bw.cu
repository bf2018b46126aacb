// write_bw.cu — microbenchmark: which store path gives the best write-only HBM bandwidth on
// B200?  (context for the paste kernel; build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/write_bw tools/write_bw.cu ; run under gpurun)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ void store16(uint4* p, uint4 v, uint64_t pol) {
    if (MODE == 0) *p = v;
    else if (MODE == 1) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == 2) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == 3) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (MODE == 4) asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    else if (MODE == 5) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) fill_kernel(uint4* out, int64_t n16, uint32_t val) {
    uint64_t pol = 0;
    if (MODE == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const uint4 v = make_uint4(val, val, val, val);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) store16<MODE>(out + i + u * stride, v, pol);
    }
    for (; i < n16; i += stride) store16<MODE>(out + i, v, pol);
}

// each CTA owns contiguous chunks (like the paste kernel's bands) instead of grid-striding
template <int MODE>
__global__ void __launch_bounds__(256) fill_chunk_kernel(uint4* out, int64_t n16, int64_t chunk16, uint32_t val) {
    const uint4 v = make_uint4(val, val, val, val);
    const int64_t nchunks = (n16 + chunk16 - 1) / chunk16;
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint4* base = out + c * chunk16;
        int64_t m = n16 - c * chunk16; if (m > chunk16) m = chunk16;
        for (int64_t i = threadIdx.x; i < m; i += 256) store16<MODE>(base + i, v, 0);
    }
}

// TMA bulk store: one elected thread per CTA streams a zeroed smem buffer to global memory
template <int BYTES>
__global__ void __launch_bounds__(128) fill_tma_kernel(unsigned char* out, int64_t nbytes) {
    __shared__ __align__(128) unsigned char zbuf[BYTES];
    for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(zbuf)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const int64_t nchunks = nbytes / BYTES;
        uint32_t src = (uint32_t)__cvta_generic_to_shared(zbuf);
        for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * BYTES), "r"(src), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <class F>
float time_it(F f, int iters = 10) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int64_t nbytes = 32ll * 100 * 512 * 1024;
    const int64_t n16 = nbytes / 16;
    unsigned char* buf;
    CK(cudaMalloc(&buf, nbytes));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, buffer %.2f GB\n", prop.name, sms, nbytes / 1e9);
    auto report = [&](const char* name, float ms) { printf("%-44s %.4f ms  %7.0f GB/s\n", name, ms, nbytes / ms / 1e6); };
    report("cudaMemsetAsync", time_it([&] { cudaMemsetAsync(buf, 0, nbytes); }));
    uint4* o = reinterpret_cast<uint4*>(buf);
    for (int per_sm : {2, 4, 8, 16}) {
        int grid = sms * per_sm;
        char nm[128];
        snprintf(nm, 128, "plain v4 unroll4 grid=%dxSM", per_sm); report(nm, time_it([&] { fill_kernel<0, 4><<<grid, 256>>>(o, n16, 0); }));
        snprintf(nm, 128, "L1::no_allocate unroll4 grid=%dxSM", per_sm); report(nm, time_it([&] { fill_kernel<1, 4><<<grid, 256>>>(o, n16, 0); }));
    }
    int grid = sms * 8;
    report("plain v4 unroll1", time_it([&] { fill_kernel<0, 1><<<grid, 256>>>(o, n16, 0); }));
    report("plain v4 unroll8", time_it([&] { fill_kernel<0, 8><<<grid, 256>>>(o, n16, 0); }));
    report("st.cs unroll4", time_it([&] { fill_kernel<2, 4><<<grid, 256>>>(o, n16, 0); }));
    report("st.cg unroll4", time_it([&] { fill_kernel<3, 4><<<grid, 256>>>(o, n16, 0); }));
    report("L2 evict_first hint unroll4", time_it([&] { fill_kernel<4, 4><<<grid, 256>>>(o, n16, 0); }));
    report("st.wt unroll4", time_it([&] { fill_kernel<5, 4><<<grid, 256>>>(o, n16, 0); }));
    report("full grid (n16/256 CTAs) plain", time_it([&] { fill_kernel<0, 1><<<(unsigned)(n16 / 256), 256>>>(o, n16, 0); }));
    for (int64_t chunk : {4096, 65536, 524288}) {
        char nm[128];
        snprintf(nm, 128, "chunked %lld B per CTA-iter, 8xSM", (long long)chunk);
        report(nm, time_it([&] { fill_chunk_kernel<1><<<grid, 256>>>(o, n16, chunk / 16, 0); }));
    }
    for (int per_sm : {1, 2, 4, 8}) {
        char nm[128];
        snprintf(nm, 128, "TMA bulk 16KB, %d CTA/SM", per_sm); report(nm, time_it([&] { fill_tma_kernel<16384><<<sms * per_sm, 128>>>(buf, nbytes); }));
        snprintf(nm, 128, "TMA bulk 32KB, %d CTA/SM", per_sm); report(nm, time_it([&] { fill_tma_kernel<32768><<<sms * per_sm, 128>>>(buf, nbytes); }));
        snprintf(nm, 128, "TMA bulk 4KB, %d CTA/SM", per_sm); report(nm, time_it([&] { fill_tma_kernel<4096><<<sms * per_sm, 128>>>(buf, nbytes); }));
    }
    // non-persistent: one CTA per contiguous chunk
    for (int64_t chunk : {8192, 16384, 32768, 65536, 131072}) {
        char nm[128];
        unsigned g = (unsigned)((nbytes + chunk - 1) / chunk);
        snprintf(nm, 128, "one CTA(256) per %lld B chunk, grid=%u", (long long)chunk, g);
        report(nm, time_it([&] { fill_chunk_kernel<1><<<g, 256>>>(o, n16, chunk / 16, 0); }));
        snprintf(nm, 128, "one CTA(256) per %lld B chunk, plain st", (long long)chunk);
        report(nm, time_it([&] { fill_chunk_kernel<0><<<g, 256>>>(o, n16, chunk / 16, 0); }));
    }
    for (int per_sm : {16, 32, 64, 128}) {
        char nm[128];
        snprintf(nm, 128, "grid-stride unroll4 grid=%dxSM", per_sm);
        report(nm, time_it([&] { fill_kernel<1, 4><<<sms * per_sm, 256>>>(o, n16, 0); }));
    }
    report("cudaMemsetAsync value 1", time_it([&] { cudaMemsetAsync(buf, 1, nbytes); }));
    report("fill val=0x01010101 16xSM", time_it([&] { fill_kernel<1, 4><<<sms * 16, 256>>>(o, n16, 0x01010101u); }));
    CK(cudaDeviceSynchronize());
    return 0;
}

// tools/microcdp.cu -- probe: a PDL-launched persistent kernel whose (never taken) failure branch repairs the call with
// a device-side TAIL launch (CUDA dynamic parallelism, cudaStreamTailLaunch).  Questions:
//   1. does the presence of the device-side launch cost anything when it is not taken (time per launch vs microcoop)?
//   2. when taken, does the tail-launched grid finish before the NEXT kernel of the stream (itself launched with
//      programmatic stream serialization) gets past griddepcontrol.wait?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true -o build/microcdp tools/microcdp.cu -lcudadevrt
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void k_repair(double* out, int n, double value)
    {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = value;
    }

__global__ void __launch_bounds__(384, 2)
    k_persist(unsigned long long* ctr, const double4* src, double4* dst, size_t n, int fail, double* out, int nout)
    {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ unsigned long long target;
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        {
        const double4 v = src[i];
        acc += v.x + v.y + v.z + v.w;
        }
    __syncthreads();
    if (threadIdx.x == 0)
        {
        const unsigned long long old = atomicAdd(ctr, 1ull);
        target = (old / gridDim.x + 1ull) * gridDim.x;
        unsigned long long v, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do
            {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            } while (v < target && t1 - t0 < 100000000ull);
        }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (fail)
        {
        // the "hand-off failed" branch: nothing is applied; one CTA asks for the repair
        if (blockIdx.x == 0 && threadIdx.x == 0)
            k_repair<<<64, 256, 0, cudaStreamTailLaunch>>>(out, nout, 42.0);
        return;
        }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = make_double4(acc, 0.0, 0.0, 0.0);
    }

// the next kernel of the stream: copies out -> seen (must observe the repaired values)
__global__ void k_consumer(const double* out, double* seen, int n)
    {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        seen[i] = out[i];
    }

int main()
    {
    const size_t n = 1000000;
    const int nout = 1 << 20;
    double4 *src, *dst;
    double *out, *seen;
    unsigned long long* ctr;
    cudaMalloc(&src, 32 * n * 8);
    cudaMalloc(&dst, 32 * n * 8);
    cudaMalloc(&ctr, 8);
    cudaMalloc(&out, 8 * nout);
    cudaMalloc(&seen, 8 * nout);
    cudaMemset(src, 0, 32 * n * 8);
    cudaMemset(ctr, 0, 8);
    cudaMemset(out, 0, 8 * nout);
    cudaMemset(seen, 0, 8 * nout);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int grid = 2 * p.multiProcessorCount;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(384);
    cfg.stream = s;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int K = 400;
    int fail = 0;
    cudaError_t first = cudaSuccess;
    for (int rep = 0; rep < 2; rep++)
        {
        cudaEventRecord(e0, s);
        for (int k = 0; k < K; k++)
            {
            const double4* sp = src + (size_t)(k % 8) * n;
            double4* dp = dst + (size_t)(k % 8) * n;
            size_t nn = n;
            int no = nout;
            void* args[] = {&ctr, &sp, &dp, &nn, &fail, &out, &no};
            cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k_persist, args);
            if (e != cudaSuccess && first == cudaSuccess)
                first = e;
            }
        cudaEventRecord(e1, s);
        cudaStreamSynchronize(s);
        }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("pdl + device-side tail launch compiled in, not taken: launch %s | sync %s | %.2f us per launch\n",
           cudaGetErrorString(first), cudaGetErrorString(cudaGetLastError()), 1e3 * ms / K);
    // taken: persistent kernel (fail = 1) then the consumer, back to back, both with the PDL attribute
    fail = 1;
        {
        const double4* sp = src;
        double4* dp = dst;
        size_t nn = n;
        int no = nout;
        void* args[] = {&ctr, &sp, &dp, &nn, &fail, &out, &no};
        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k_persist, args);
        cudaLaunchConfig_t c2 = cfg;
        c2.gridDim = dim3(256);
        c2.blockDim = dim3(256);
        const double* o = out;
        void* a2[] = {&o, &seen, &no};
        cudaError_t e2 = cudaLaunchKernelExC(&c2, (const void*)k_consumer, a2);
        cudaError_t e3 = cudaStreamSynchronize(s);
        double* host = (double*)malloc(8 * nout);
        cudaMemcpy(host, seen, 8 * nout, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < nout; i++)
            bad += host[i] != 42.0;
        printf("taken: launch %s / consumer %s / sync %s | consumer saw the repaired values in %d of %d slots\n",
               cudaGetErrorString(e), cudaGetErrorString(e2), cudaGetErrorString(e3), nout - bad, nout);
        }
    return 0;
    }

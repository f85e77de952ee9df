/*
 * ehyb_solver.cu -- (preconditioned) conjugate gradients on top of the EHYB product
 * (SURVEY.md section 8f-4).
 *
 * The reference carries the skeleton of this solver and never calls it: kernelInitializeAll /
 * kernelInitializeR / kernelMyxpy and their launchers (kernel.cu:13-42, :288-321: r = b,
 * p = z, y = x + gamma*y), matrixCOO.diag (the Jacobi preconditioner, filled by the reader),
 * cb_s.PRECOND and the never-written realIter of spmvGPuEHYB.  Here the loop is complete:
 *
 *     x = 0, r = b, z = D^-1 r, p = z
 *     repeat:  q = A p (ehyb_spmv);  alpha = (r.z)/(p.q);  x += alpha p;  r -= alpha q;
 *              z = D^-1 r;  beta = (r.z)_new/(r.z);  p = z + beta p
 *
 * with every scalar kept ON THE DEVICE (slots indexed by iteration parity, reductions by warp
 * shuffles + one atomicAdd per block), so an iteration is the product plus three small fused
 * kernels on the session stream and never waits for the host; the host looks at |r|^2 only every
 * `check_every` iterations.  Vectors use the permuted numbering of the session.  Everything
 * goes through the public C ABI of the session (ehyb_spmv, ehyb_stream): no kernel internals.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../host/common.h"

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            rc = ehyb_fail(EHYB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            goto done;                                                                               \
        }                                                                                            \
    } while (0)

namespace {

/* device scalars: RZ[2] (r.z, by parity), PQ[2] (p.q), RR[2] (r.r), BB */
enum { S_RZ = 0, S_PQ = 2, S_RR = 4, S_BB = 6, S_COUNT = 8 };

constexpr int kThreads = 256;

__device__ __forceinline__ void block_add(double v, double *slot)
{
    __shared__ double part[kThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < kThreads / 32 ? part[lane] : 0.0;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) atomicAdd(slot, v);
    }
    __syncthreads();
}

/* x = 0, r = b, z = dinv*r (or r), p = z;  RZ[0] = r.z, RR[0] = r.r, BB = b.b */
__global__ void __launch_bounds__(kThreads) pcg_init(int n, const double *__restrict__ b, const double *__restrict__ dinv,
                                                       double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                                                       double *__restrict__ p, double *__restrict__ s)
{
    double rz = 0.0, rr = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double bi = b[i], zi = dinv ? dinv[i] * bi : bi;
        x[i] = 0.0; r[i] = bi; z[i] = zi; p[i] = zi;
        rz += bi * zi; rr += bi * bi;
    }
    block_add(rz, s + S_RZ);
    block_add(rr, s + S_RR);
    block_add(rr, s + S_BB);
}

/* PQ[k] = p.q; also clears the slots iteration k+1 will accumulate into */
__global__ void __launch_bounds__(kThreads) pcg_dot_pq(int n, const double *__restrict__ p, const double *__restrict__ q, double *__restrict__ s, int k)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { /* nobody reads these before the next kernel of this iteration */
        s[S_RZ + ((k + 1) & 1)] = 0.0;
        s[S_RR + ((k + 1) & 1)] = 0.0;
        s[S_PQ + ((k + 1) & 1)] = 0.0;
    }
    double v = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += p[i] * q[i];
    block_add(v, s + S_PQ + (k & 1));
}

/* alpha = RZ[k]/PQ[k];  x += alpha p;  r -= alpha q;  z = dinv*r;  RZ[k+1] = r.z, RR[k+1] = r.r */
__global__ void __launch_bounds__(kThreads) pcg_update_xr(int n, const double *__restrict__ p, const double *__restrict__ q,
                                                            const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r,
                                                            double *__restrict__ z, double *__restrict__ s, int k)
{
    const double pq = s[S_PQ + (k & 1)];
    const double alpha = pq != 0.0 ? s[S_RZ + (k & 1)] / pq : 0.0;
    double rz = 0.0, rr = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        const double zi = dinv ? dinv[i] * ri : ri;
        r[i] = ri; z[i] = zi;
        rz += ri * zi; rr += ri * ri;
    }
    block_add(rz, s + S_RZ + ((k + 1) & 1));
    block_add(rr, s + S_RR + ((k + 1) & 1));
}

/* beta = RZ[k+1]/RZ[k];  p = z + beta p   (the reference's myxpy: y = x + gamma*y, kernel.cu:287-296) */
__global__ void __launch_bounds__(kThreads) pcg_update_p(int n, const double *__restrict__ z, double *__restrict__ p, const double *__restrict__ s, int k)
{
    const double rz = s[S_RZ + (k & 1)];
    const double beta = rz != 0.0 ? s[S_RZ + ((k + 1) & 1)] / rz : 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}

/* |b - q|^2 with q = A x: the true residual at the end */
__global__ void __launch_bounds__(kThreads) pcg_true_residual(int n, const double *__restrict__ b, const double *__restrict__ q, double *__restrict__ out)
{
    double v = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double d = b[i] - q[i];
        v += d * d;
    }
    block_add(v, out);
}

} /* namespace */

extern "C" void ehyb_pcg_opts_default(ehyb_pcg_opts *o)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->max_iters = 1000;
    o->rtol = 1e-10;
    o->check_every = 8;
}

extern "C" int ehyb_pcg_solve(ehyb_handle *h, const double *diag_h, const double *b_h, double *x_h, const ehyb_pcg_opts *opts,
                              ehyb_pcg_result *res)
{
    if (!h || !b_h || !x_h || !res) return ehyb_fail(EHYB_ERR_ARG, "ehyb_pcg_solve: NULL argument");
    ehyb_pcg_opts o;
    if (opts) o = *opts;
    else ehyb_pcg_opts_default(&o);
    if (o.max_iters <= 0 || !(o.rtol >= 0.0)) return ehyb_fail(EHYB_ERR_ARG, "ehyb_pcg_solve: bad options");
    if (o.check_every <= 0) o.check_every = 8;
    memset(res, 0, sizeof *res);
    int rc = EHYB_OK;
    int64_t n64 = 0, ncols = 0;
    rc = ehyb_session_size(h, &n64, &ncols);
    if (rc) return rc;
    if (ncols != n64) return ehyb_fail(EHYB_ERR_ARG, "ehyb_pcg_solve: the session is a distributed block (halo columns)");
    const int n = (int)n64;
    cudaStream_t st = (cudaStream_t)ehyb_stream(h);
    double *b = NULL, *x = NULL, *r = NULL, *z = NULL, *p = NULL, *q = NULL, *dinv = NULL, *s = NULL, *tmp = NULL, *host = NULL;
    cudaEvent_t e0 = NULL, e1 = NULL;
    int dev = 0, sms = 0, k = 0;
    float ms = 0.f;
    double sc[S_COUNT];
    const size_t vb = sizeof(double) * ((size_t)n + 2);
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    {
        const int grid = sms * 4;
        CU(cudaMalloc(&b, vb)); CU(cudaMalloc(&x, vb)); CU(cudaMalloc(&r, vb)); CU(cudaMalloc(&z, vb));
        CU(cudaMalloc(&p, vb)); CU(cudaMalloc(&q, vb)); CU(cudaMalloc(&s, sizeof(double) * S_COUNT)); CU(cudaMalloc(&tmp, sizeof(double)));
        CU(cudaMallocHost(&host, sizeof(double) * S_COUNT));
        CU(cudaMemcpyAsync(b, b_h, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
        if (diag_h) {
            /* D^-1 on the host once (a zero diagonal entry leaves the row unpreconditioned) */
            double *inv = (double *)malloc(sizeof(double) * (size_t)n);
            if (!inv) { rc = ehyb_fail(EHYB_ERR_NOMEM, "ehyb_pcg_solve: out of memory"); goto done; }
            for (int i = 0; i < n; ++i) inv[i] = diag_h[i] != 0.0 ? 1.0 / diag_h[i] : 1.0;
            cudaError_t e = cudaMalloc(&dinv, vb);
            if (e == cudaSuccess) e = cudaMemcpy(dinv, inv, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
            free(inv);
            if (e != cudaSuccess) { rc = ehyb_fail(EHYB_ERR_CUDA, "ehyb_pcg_solve: %s", cudaGetErrorString(e)); goto done; }
        }
        CU(cudaMemsetAsync(s, 0, sizeof(double) * S_COUNT, st));
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, st));
        pcg_init<<<grid, kThreads, 0, st>>>(n, b, dinv, x, r, z, p, s);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(host, s, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const double bb = host[S_BB];
        res->rel_residual = bb > 0.0 ? 1.0 : 0.0;
        if (bb == 0.0) { res->converged = 1; }
        while (bb > 0.0 && k < o.max_iters && !res->converged) {
            const int until = k + o.check_every < o.max_iters ? k + o.check_every : o.max_iters;
            for (; k < until; ++k) {
                rc = ehyb_spmv(h, p, q); /* q = A p, asynchronous on the session stream */
                if (rc) goto done;
                pcg_dot_pq<<<grid, kThreads, 0, st>>>(n, p, q, s, k);
                pcg_update_xr<<<grid, kThreads, 0, st>>>(n, p, q, dinv, x, r, z, s, k);
                pcg_update_p<<<grid, kThreads, 0, st>>>(n, z, p, s, k);
            }
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(host, s, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            const double rr = host[S_RR + (k & 1)];
            res->rel_residual = sqrt(rr / bb);
            if (!(rr == rr)) { rc = ehyb_fail(EHYB_ERR_ARG, "ehyb_pcg_solve: breakdown (NaN residual) at iteration %d: is the matrix symmetric positive definite?", k); goto done; }
            if (res->rel_residual <= o.rtol) res->converged = 1;
        }
        CU(cudaEventRecord(e1, st));
        /* true residual |b - A x| / |b| */
        rc = ehyb_spmv(h, x, q);
        if (rc) goto done;
        CU(cudaMemsetAsync(tmp, 0, sizeof(double), st));
        pcg_true_residual<<<grid, kThreads, 0, st>>>(n, b, q, tmp);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(sc, tmp, sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(x_h, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        res->iters = k;
        res->ms = ms;
        res->true_rel_residual = bb > 0.0 ? sqrt(sc[0] / bb) : 0.0;
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(b); cudaFree(x); cudaFree(r); cudaFree(z); cudaFree(p); cudaFree(q); cudaFree(dinv); cudaFree(s); cudaFree(tmp);
    if (host) cudaFreeHost(host);
    return rc;
}

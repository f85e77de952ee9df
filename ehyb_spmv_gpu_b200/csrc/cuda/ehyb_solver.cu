/*
 * ehyb_solver.cu -- (preconditioned) conjugate gradients on top of the EHYB product, on one GPU
 * (ehyb_pcg_solve) and over the row blocks of several (ehyb_mg_pcg_solve): SURVEY.md section 8f-4.
 *
 * The reference carries the skeleton of this solver and never calls it: kernelInitializeAll /
 * kernelInitializeR / kernelMyxpy and their launchers (kernel.cu:13-42, :288-321: r = b,
 * p = z, y = x + gamma*y), matrixCOO.diag (the Jacobi preconditioner, filled by the reader),
 * cb_s.PRECOND and the never-written realIter of spmvGPuEHYB.  Here the loop is complete:
 *
 *     x = 0, r = b, z = D^-1 r, p = z
 *     repeat:  q = A p;  alpha = (r.z)/(p.q);  x += alpha p;  r -= alpha q;
 *              z = D^-1 r;  beta = (r.z)_new/(r.z);  p = z + beta p
 *
 * An iteration is THREE launches on the session stream, chained by programmatic dependent launch,
 * and never waits for the host (it looks at |r|^2 every `check_every` iterations):
 *   1. the product WITH p.q inside (ehyb_spmv_dot: p is the kernel's x window in shared memory, so
 *      y[r] * p[r] costs no memory traffic) - round 1 read p and q again in a kernel of its own;
 *   2. pcg_update_xr: alpha from the device scalars, x, r, z and the partial sums of r.z and r.r;
 *   3. pcg_update_p.
 * All scalars live ON THE DEVICE in slots that rotate with the iteration (three {r.z, r.r} pairs,
 * two p.q): the slot an iteration accumulates into was zeroed one iteration earlier by a kernel
 * that ran strictly between its last reader and its next writer.
 *
 * Several GPUs: every rank owns the rows of its block of every vector; the product is the session's
 * distributed product (halo exchange inside the kernel), and after launch 1 and launch 2 the
 * partial sums are summed over the GPUs by ehyb_mg_allreduce_sum (peer-memory mailbox, rank order:
 * every rank holds the same bits, so every rank takes the same decisions).
 *
 * Vectors use the permuted numbering of the session.  Everything goes through the public C ABI of
 * the sessions (ehyb_spmv[_dot], ehyb_mg_spmv[_dot], ehyb_mg_allreduce_sum, ehyb_stream).
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

/* device scalars: G[j] = {r.z, r.r} for j = iteration mod 3 at [2j, 2j+1]; PQ[k & 1] at 6, 7; the
 * true residual at 8 */
enum { S_G = 0, S_PQ = 6, S_TRUE = 8, S_COUNT = 16 };

constexpr int kThreads = 256;

__device__ __forceinline__ void pdl_prologue()
{
    /* the next kernel of the stream may be scheduled as this grid's blocks retire; this grid's own
     * inputs belong to its predecessor */
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void block_add2(double a, double b, double *slot)
{
    __shared__ double part[2][kThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, off);
        b += __shfl_down_sync(0xffffffffu, b, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { part[0][warp] = a; part[1][warp] = b; }
    __syncthreads();
    if (warp == 0) {
        a = lane < kThreads / 32 ? part[0][lane] : 0.0;
        b = lane < kThreads / 32 ? part[1][lane] : 0.0;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, off);
            b += __shfl_down_sync(0xffffffffu, b, off);
        }
        if (lane == 0) { atomicAdd(slot, a); atomicAdd(slot + 1, b); }
    }
}

__device__ __forceinline__ void block_add1(double a, double *slot)
{
    __shared__ double part1[kThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part1[warp] = a;
    __syncthreads();
    if (warp == 0) {
        a = lane < kThreads / 32 ? part1[lane] : 0.0;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
        if (lane == 0) atomicAdd(slot, a);
    }
}

/* x = 0, r = b, z = dinv*r (or r), p = z;  G[0] = {r.z, r.r} */
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
    block_add2(rz, rr, s + S_G);
}

/* PQ[k & 1] += p.q - only where the product cannot carry the dot itself (ehyb_spmv_dot_supported) */
__global__ void __launch_bounds__(kThreads) pcg_dot_pq(int n, const double *__restrict__ p, const double *__restrict__ q, double *__restrict__ s, int k)
{
    pdl_prologue();
    double v = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += p[i] * q[i];
    block_add1(v, s + S_PQ + (k & 1));
}

/* alpha = G[k%3].rz / PQ[k&1];  x += alpha p;  r -= alpha q;  z = dinv*r;  G[(k+1)%3] += {r.z, r.r};
 * zeroes the slots the NEXT iteration accumulates into */
__global__ void __launch_bounds__(kThreads) pcg_update_xr(int n, const double *__restrict__ p, const double *__restrict__ q,
                                                            const double *__restrict__ dinv, double *__restrict__ x, double *__restrict__ r,
                                                            double *__restrict__ z, double *__restrict__ s, int k)
{
    pdl_prologue();
    const double pq = s[S_PQ + (k & 1)];
    const double alpha = pq != 0.0 ? s[S_G + 2 * (k % 3)] / pq : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        /* G[(k+2)%3]: last read by update_p of iteration k-1, next written by update_xr of k+1;
         * PQ[(k+1)&1]: last read by update_xr of k-1, next written by the product of k+1 */
        s[S_G + 2 * ((k + 2) % 3)] = 0.0;
        s[S_G + 2 * ((k + 2) % 3) + 1] = 0.0;
        s[S_PQ + ((k + 1) & 1)] = 0.0;
    }
    double rz = 0.0, rr = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        const double zi = dinv ? dinv[i] * ri : ri;
        r[i] = ri; z[i] = zi;
        rz += ri * zi; rr += ri * ri;
    }
    block_add2(rz, rr, s + S_G + 2 * ((k + 1) % 3));
}

/* beta = G[(k+1)%3].rz / G[k%3].rz;  p = z + beta p   (the reference's myxpy: y = x + gamma*y, kernel.cu:287-296) */
__global__ void __launch_bounds__(kThreads) pcg_update_p(int n, const double *__restrict__ z, double *__restrict__ p, const double *__restrict__ s, int k)
{
    pdl_prologue();
    const double rz = s[S_G + 2 * (k % 3)];
    const double beta = rz != 0.0 ? s[S_G + 2 * ((k + 1) % 3)] / rz : 0.0;
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
    block_add1(v, out);
}

/* programmatic dependent launch of a vector kernel on the session stream */
template <typename... Params, typename... Args>
cudaError_t launch_pdl(void (*kernel)(Params...), int grid, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

/* the loop, on one session (mg == NULL) or on this rank's block of a distributed one */
int pcg_run(ehyb_handle *h, ehyb_mg_session *mg, const double *diag_h, const double *b_h, double *x_h, const ehyb_pcg_opts *opts,
            ehyb_pcg_result *res, const char *who)
{
    if (!h || !b_h || !x_h || !res) return ehyb_fail(EHYB_ERR_ARG, "%s: NULL argument", who);
    ehyb_pcg_opts o;
    if (opts) o = *opts;
    else ehyb_pcg_opts_default(&o);
    if (o.max_iters <= 0 || !(o.rtol >= 0.0)) return ehyb_fail(EHYB_ERR_ARG, "%s: bad options", who);
    if (o.check_every <= 0) o.check_every = 8;
    memset(res, 0, sizeof *res);
    int rc = EHYB_OK;
    int64_t n64 = 0, ncols = 0;
    rc = ehyb_session_size(h, &n64, &ncols);
    if (rc) return rc;
    if (!mg && ncols != n64) return ehyb_fail(EHYB_ERR_ARG, "%s: the session is a distributed block (halo columns): use ehyb_mg_pcg_solve", who);
    const int n = (int)n64;
    int nranks = 1;
    if (mg) ehyb_mg_session_ranks(mg, NULL, &nranks);
    cudaStream_t st = (cudaStream_t)ehyb_stream(h);
    double *b = NULL, *x = NULL, *r = NULL, *z = NULL, *p = NULL, *q = NULL, *dinv = NULL, *s = NULL, *host = NULL;
    cudaEvent_t e0 = NULL, e1 = NULL;
    int prevDev = -1, sms = 0, k = 0;
    float ms = 0.f;
    const int dev = ehyb_session_device(h);
    const size_t vb = sizeof(double) * ((size_t)n + 2);
    const bool fused = mg ? ehyb_mg_spmv_dot_supported(mg) != 0 : ehyb_spmv_dot_supported(h) != 0;
    const bool pdl = getenv("EHYB_PDL") == NULL || atoi(getenv("EHYB_PDL")) != 0;
    /* the session may live on another GPU than the caller's current one */
    cudaGetDevice(&prevDev);
    CU(cudaSetDevice(dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    {
        const int grid = sms * 8;
        auto product = [&](double *xin, double *yout, double *dot) -> int {
            if (mg) return dot ? ehyb_mg_spmv_dot(mg, xin, yout, dot) : ehyb_mg_spmv(mg, xin, yout);
            return dot ? ehyb_spmv_dot(h, xin, yout, dot) : ehyb_spmv(h, xin, yout);
        };
        auto peers_ok = [&]() -> int { /* a rank that did not show up within the time limit */
            int to = 0;
            if (!mg) return EHYB_OK;
            int rc2 = ehyb_mg_status(mg, &to);
            if (rc2) return rc2;
            return to ? ehyb_fail(EHYB_ERR_PEER, "%s: a neighbour GPU did not answer within the peer time limit", who) : EHYB_OK;
        };
        CU(cudaMalloc(&b, vb)); CU(cudaMalloc(&x, vb)); CU(cudaMalloc(&r, vb)); CU(cudaMalloc(&z, vb));
        CU(cudaMalloc(&p, sizeof(double) * ((size_t)ncols + 2))); /* a distributed product may keep the halo behind the own entries */
        CU(cudaMalloc(&q, vb)); CU(cudaMalloc(&s, sizeof(double) * S_COUNT));
        CU(cudaMallocHost(&host, sizeof(double) * S_COUNT));
        CU(cudaMemsetAsync(p, 0, sizeof(double) * ((size_t)ncols + 2), st));
        CU(cudaMemcpyAsync(b, b_h, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
        if (diag_h) {
            /* D^-1 on the host once (a zero diagonal entry leaves the row unpreconditioned) */
            double *inv = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
            if (!inv) { rc = ehyb_fail(EHYB_ERR_NOMEM, "%s: out of memory", who); goto done; }
            for (int i = 0; i < n; ++i) inv[i] = diag_h[i] != 0.0 ? 1.0 / diag_h[i] : 1.0;
            cudaError_t e = cudaMalloc(&dinv, vb);
            if (e == cudaSuccess) e = cudaMemcpy(dinv, inv, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
            free(inv);
            if (e != cudaSuccess) { rc = ehyb_fail(EHYB_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e)); goto done; }
        }
        CU(cudaMemsetAsync(s, 0, sizeof(double) * S_COUNT, st));
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, st));
        pcg_init<<<grid, kThreads, 0, st>>>(n, b, dinv, x, r, z, p, s);
        CU(cudaGetLastError());
        if (mg) { rc = ehyb_mg_allreduce_sum(mg, s + S_G, 2); if (rc) goto done; }
        CU(cudaMemcpyAsync(host, s, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        rc = peers_ok();
        if (rc) goto done;
        const double bb = host[S_G + 1]; /* x0 = 0: r0 = b */
        res->rel_residual = bb > 0.0 ? 1.0 : 0.0;
        if (bb == 0.0) { res->converged = 1; }
        while (bb > 0.0 && k < o.max_iters && !res->converged) {
            const int until = k + o.check_every < o.max_iters ? k + o.check_every : o.max_iters;
            for (; k < until; ++k) {
                double *pq = s + S_PQ + (k & 1);
                rc = product(p, q, fused ? pq : NULL); /* q = A p (+ p.q), asynchronous on the session stream */
                if (rc) goto done;
                if (!fused) CU(launch_pdl(pcg_dot_pq, grid, st, pdl, n, (const double *)p, (const double *)q, s, k));
                if (mg) { rc = ehyb_mg_allreduce_sum(mg, pq, 1); if (rc) goto done; }
                CU(launch_pdl(pcg_update_xr, grid, st, pdl, n, (const double *)p, (const double *)q, (const double *)dinv, x, r, z, s, k));
                if (mg) { rc = ehyb_mg_allreduce_sum(mg, s + S_G + 2 * ((k + 1) % 3), 2); if (rc) goto done; }
                CU(launch_pdl(pcg_update_p, grid, st, pdl, n, (const double *)z, p, (const double *)s, k));
            }
            CU(cudaMemcpyAsync(host, s, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            rc = peers_ok();
            if (rc) goto done;
            const double rr = host[S_G + 2 * (k % 3) + 1];
            res->rel_residual = sqrt(rr / bb);
            if (!(rr == rr)) { rc = ehyb_fail(EHYB_ERR_ARG, "%s: breakdown (NaN residual) at iteration %d: is the matrix symmetric positive definite?", who, k); goto done; }
            if (res->rel_residual <= o.rtol) res->converged = 1;
        }
        CU(cudaEventRecord(e1, st));
        /* true residual |b - A x| / |b| (a distributed product reads its x with halo room behind: p's buffer) */
        CU(cudaMemcpyAsync(p, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        rc = product(p, q, NULL);
        if (rc) goto done;
        CU(cudaMemsetAsync(s + S_TRUE, 0, 2 * sizeof(double), st));
        pcg_true_residual<<<grid, kThreads, 0, st>>>(n, b, q, s + S_TRUE);
        CU(cudaGetLastError());
        if (mg) { rc = ehyb_mg_allreduce_sum(mg, s + S_TRUE, 1); if (rc) goto done; }
        CU(cudaMemcpyAsync(host, s, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(x_h, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        rc = peers_ok();
        if (rc) goto done;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        res->iters = k;
        res->ms = ms;
        res->true_rel_residual = bb > 0.0 ? sqrt(host[S_TRUE] / bb) : 0.0;
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(b); cudaFree(x); cudaFree(r); cudaFree(z); cudaFree(p); cudaFree(q); cudaFree(dinv); cudaFree(s);
    if (host) cudaFreeHost(host);
    if (prevDev >= 0) cudaSetDevice(prevDev);
    return rc;
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
    return pcg_run(h, NULL, diag_h, b_h, x_h, opts, res, "ehyb_pcg_solve");
}

extern "C" int ehyb_mg_pcg_solve(ehyb_mg_session *s, const double *diag_h, const double *b_h, double *x_h, const ehyb_pcg_opts *opts,
                                 ehyb_pcg_result *res)
{
    if (!s) return ehyb_fail(EHYB_ERR_ARG, "ehyb_mg_pcg_solve: NULL session");
    ehyb_handle *h = NULL;
    int rc = ehyb_mg_session_handle(s, &h);
    if (rc) return rc;
    return pcg_run(h, s, diag_h, b_h, x_h, opts, res, "ehyb_mg_pcg_solve");
}

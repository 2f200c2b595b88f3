// selftest.cu -- on-device check that the branch-free division / square root of common.cuh return the same bits
// as nvcc's IEEE div.rn.f64 / sqrt.rn.f64 inside their guaranteed operand range (and raise the flag outside).
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long &x)
{
    x += 0x9E3779B97F4A7C15ULL;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// random double with a uniformly random mantissa/sign and an exponent in [1023-span, 1023+span]
__device__ __forceinline__ double random_double(unsigned long long &state, int span)
{
    const unsigned long long r = splitmix64(state);
    const unsigned long long mant = r & 0x000FFFFFFFFFFFFFULL;
    const unsigned long long sign = r & 0x8000000000000000ULL;
    const long long e = 1023 - span + (long long)((r >> 52) & 0x7FF) % (2 * span + 1);
    return __longlong_as_double((long long)(sign | ((unsigned long long)e << 52) | mant));
}

__global__ void k_selftest(unsigned long long seed, int per_thread, unsigned long long *mismatch)
{
    unsigned long long state = seed + 0x1234567ULL * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad_div = 0, bad_sqrt = 0, bad_flag = 0, bad_shared = 0;
    for (int it = 0; it < per_thread; it++) {
        // (1) operands spread over the whole guaranteed range; (2) operands close to each other (quotients near 1,
        // the hard rounding cases); (3) exact and near-exact quotients
        const int mode = it & 3;
        double a, b;
        if (mode == 0) { a = random_double(state, 900); b = random_double(state, 120); }
        else if (mode == 1) { a = random_double(state, 3); b = random_double(state, 3); }
        else if (mode == 2) {
            b = random_double(state, 40);
            const double q = (double)(long long)(splitmix64(state) & 0xFFFFF) + 1.0;
            a = __dmul_rn(b, q);                                  // a/b within an ulp of an integer
        } else { a = (it & 4) ? 0.0 : random_double(state, 200); b = random_double(state, 100); }

        RangeFlag f;
        const double q = div_rn_flagged(a, b, f);
        const double q_ref = __ddiv_rn(a, b);
        if (f.bad()) bad_flag++;
        if (__double_as_longlong(q) != __double_as_longlong(q_ref)) bad_div++;

        // shared reciprocal: second quotient with the same divisor
        const double a2 = random_double(state, 100);
        RangeFlag f2;
        const double r = rcp_refined(b);
        if (__double_as_longlong(div_with_rcp(a2, b, r)) != __double_as_longlong(__ddiv_rn(a2, b))) bad_shared++;

        const double s_in = fabs(mode == 0 ? random_double(state, 900) : a);
        RangeFlag f3;
        const double s = sqrt_rn_flagged(s_in, f3);
        if (f3.bad()) bad_flag++;
        if (__double_as_longlong(s) != __double_as_longlong(__dsqrt_rn(s_in))) bad_sqrt++;
        (void)f2;
    }
    // out-of-range operands must raise the flag
    unsigned long long missed = 0;
    {
        RangeFlag f;
        div_rn_flagged(1.0, 0.0, f);                       if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(1.0, 1e-200, f);   if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(1.0, 1e+200, f);   if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(1e-300, 1.0, f);   if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(4.9e-324, 1.0, f); if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(1e+300, 1.0, f);   if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(__longlong_as_double(0x7FF8000000000000LL), 1.0, f); if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(1.0, __longlong_as_double(0x7FF0000000000000LL), f); if (!f.bad()) missed++;
        f = RangeFlag(); sqrt_rn_flagged(-1.0, f);         if (!f.bad()) missed++;
        f = RangeFlag(); sqrt_rn_flagged(1e-305, f);       if (!f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(0.0, 3.0, f);      if (f.bad()) missed++;     // zero dividend is in range
        f = RangeFlag(); div_rn_flagged(-0.0, 3.0, f);     if (f.bad()) missed++;
        f = RangeFlag(); div_rn_flagged(1e-270, 1e30, f);  if (f.bad()) missed++;     // tiny dividend, still in range
        f = RangeFlag(); sqrt_rn_flagged(0.0, f);          if (f.bad()) missed++;
    }
    atomicAdd(&mismatch[0], bad_div);
    atomicAdd(&mismatch[1], bad_sqrt);
    atomicAdd(&mismatch[2], bad_shared);
    atomicAdd(&mismatch[3], bad_flag);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&mismatch[4], missed);
}

}   // namespace

extern "C" int armon_selftest_math(armon_ctx *ctx, uint64_t seed, uint64_t n_samples, uint64_t mismatch[5])
{
    ARMON_CHECK_ARG(ctx && mismatch, "null argument");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    unsigned long long *d = reinterpret_cast<unsigned long long *>(ctx->scratch);
    ARMON_CUDA(cudaMemsetAsync(d, 0, 5 * sizeof(unsigned long long), ctx->stream));
    const int threads = 256, blocks = 148 * 8;
    const int per_thread = (int)((n_samples + (uint64_t)threads * blocks - 1) / ((uint64_t)threads * blocks));
    k_selftest<<<blocks, threads, 0, ctx->stream>>>(seed, per_thread < 1 ? 1 : per_thread, d);
    ARMON_LAUNCH_CHECK(ctx);
    unsigned long long *h = reinterpret_cast<unsigned long long *>(ctx->pinned);
    ARMON_CUDA(cudaMemcpyAsync(h, d, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 5; k++) mismatch[k] = h[k];
    return ARMON_OK;
}

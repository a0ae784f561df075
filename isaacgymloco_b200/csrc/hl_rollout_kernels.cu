// GAE / returns (sm_100a).  Reference: HIMRolloutStorage.compute_returns,
// rsl_rl/rsl_rl/storage/him_rollout_storage.py:113-127 (== amp_rollout_storage.py:141-155).
#include "hl_common.cuh"

// One thread per env walks its T transitions backwards; consecutive threads touch consecutive
// addresses of every (T,N,1) tensor, so all traffic is coalesced.  Each op is rounded on its own
// in the reference's order, which makes `returns` bit-identical to eager torch:
//   nnt = 1 - done;  delta = (r + (nnt*gamma)*V_next) - V;  A = delta + ((nnt*gamma)*lam)*A;
//   ret = A + V
// Raw advantages (ret - V) are written and their moments accumulated in float64 for the
// normalisation pass (all-reduced first when envs are sharded over GPUs).
template <int UNROLL>
__global__ void __launch_bounds__(256) hl_gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                          const uint8_t* __restrict__ dones,
                                                          const float* __restrict__ last_values,
                                                          float* __restrict__ returns, float* __restrict__ adv,
                                                          double* __restrict__ moments, int t_len, long long n,
                                                          float gamma, float lam) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (e < n) {
    float next_v = last_values[e];
    float a = 0.0f;
    int t = t_len - 1;
    for (; t - (UNROLL - 1) >= 0; t -= UNROLL) {
      float r[UNROLL], v[UNROLL], d[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {  // independent loads first: UNROLL requests in flight
        const long long i = (long long)(t - u) * n + e;
        r[u] = __ldg(rewards + i);
        v[u] = __ldg(values + i);
        d[u] = (float)__ldg(dones + i);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long i = (long long)(t - u) * n + e;
        const float g = __fmul_rn(__fsub_rn(1.0f, d[u]), gamma);
        const float delta = __fsub_rn(__fadd_rn(r[u], __fmul_rn(g, next_v)), v[u]);
        a = __fadd_rn(delta, __fmul_rn(__fmul_rn(g, lam), a));
        const float ret = __fadd_rn(a, v[u]);
        returns[i] = ret;
        const float ad = __fsub_rn(ret, v[u]);
        adv[i] = ad;
        s1 += (double)ad;
        s2 += (double)ad * (double)ad;
        next_v = v[u];
      }
    }
    for (; t >= 0; --t) {
      const long long i = (long long)t * n + e;
      const float rr = __ldg(rewards + i), vv = __ldg(values + i), dd = (float)__ldg(dones + i);
      const float g = __fmul_rn(__fsub_rn(1.0f, dd), gamma);
      const float delta = __fsub_rn(__fadd_rn(rr, __fmul_rn(g, next_v)), vv);
      a = __fadd_rn(delta, __fmul_rn(__fmul_rn(g, lam), a));
      const float ret = __fadd_rn(a, vv);
      returns[i] = ret;
      const float ad = __fsub_rn(ret, vv);
      adv[i] = ad;
      s1 += (double)ad;
      s2 += (double)ad * (double)ad;
      next_v = vv;
    }
  }
  // block reduction of the two moments: shuffle inside warps, shared memory across warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  __shared__ double w1[8], w2[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { w1[wid] = s1; w2[wid] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a1 = 0.0, a2 = 0.0;
    for (int w = 0; w < 8; ++w) { a1 += w1[w]; a2 += w2[w]; }
    atomicAdd(moments + 0, a1);
    atomicAdd(moments + 1, a2);
    if (blockIdx.x == 0) atomicAdd(moments + 2, (double)t_len * (double)n);
  }
}

extern "C" int hl_gae_scan(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                           float* returns, float* advantages, double* moments, int32_t t_len, int64_t n, float gamma,
                           float lam, void* stream) {
  HL_CHECK_ARG(rewards && values && dones && last_values && returns && advantages && moments, "null pointer");
  HL_CHECK_ARG(t_len >= 0 && n >= 0, "negative size");
  if (n == 0 || t_len == 0) return HL_OK;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  hl_gae_scan_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(rewards, values, dones, last_values, returns, advantages,
                                                                moments, t_len, n, gamma, lam);
  HL_CHECK_LAUNCH();
  return HL_OK;
}

// advantages = (adv - mean) / (std + 1e-8), std unbiased (torch default) -- :126-127
__global__ void __launch_bounds__(256) hl_adv_normalize_kernel(float* __restrict__ adv, const double* __restrict__ moments,
                                                               long long n) {
  const double cnt = moments[2];
  const double mean = moments[0] / cnt;
  double var = (moments[1] - moments[0] * mean) / (cnt - 1.0);
  var = var > 0.0 ? var : 0.0;
  const float meanf = (float)mean;
  const float denom = (float)sqrt(var) + 1e-8f;
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n && (((uintptr_t)adv & 15) == 0)) {
    float4 v = *reinterpret_cast<float4*>(adv + i4);
    v.x = (v.x - meanf) / denom;
    v.y = (v.y - meanf) / denom;
    v.z = (v.z - meanf) / denom;
    v.w = (v.w - meanf) / denom;
    *reinterpret_cast<float4*>(adv + i4) = v;
  } else {
    for (long long i = i4; i < n && i < i4 + 4; ++i) adv[i] = (adv[i] - meanf) / denom;
  }
}

extern "C" int hl_adv_normalize(float* advantages, const double* moments, int64_t n_elems, void* stream) {
  HL_CHECK_ARG(advantages && moments, "null pointer");
  if (n_elems <= 0) return HL_OK;
  const long long threads = (n_elems + 3) / 4;
  hl_adv_normalize_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(advantages, moments, n_elems);
  HL_CHECK_LAUNCH();
  return HL_OK;
}

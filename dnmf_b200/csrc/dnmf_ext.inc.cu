// North-star EXTENSION (no counterpart in the reference, default OFF): gradients of the loss with
// respect to the SHARED parameters -- neuron positions pos[K][3], widths sigma[K] and a scalar background
// b -- so that they can be learned next to the per-frame deformation.  Frames are sharded over GPUs, so
// these are the only gradients that need a collective (one small all-reduce per iteration).
//
//   Yhat_t(p) = sum_k C[k,t] A_t(p,k; pos, sigma) + b,     A = discretised Gaussian resampled trilinearly
//   dL/dpos_kd  = (2/BN) sum_t sum_p r_t(p) C[k,t] a'_kd prod_{e!=d} a_ke,  a' = lerp of dG/dpos
//   dL/dsigma_k = (2/BN) sum_t sum_p r_t(p) C[k,t] sum_d a^s_kd prod_{e!=d} a_ke,  a^s = lerp of dG/dsigma
//   dL/db       = (2/BN) sum_t sum_p r_t(p)
//
// fit_tile_kernel<MODE=2> writes the residual r; param_grad_kernel (one warp per 8x8xtz tile-frame) stages the
// residual tile and the listed neurons' table slices (G, D) in shared memory, eight slots at a time, and forms the
// lerps of dG/dpos and dG/dsigma from the SAME (G, D) entries -- for G_i = exp(-d_i^2 / sigma^2), d_i = i - pos:
//   lerp(dG/dpos)   = (2 / sigma^2) (d_i a + f G_{i+1}),        a = G_i + f D_i,  G_{i+1} = G_i + D_i
//   lerp(dG/dsigma) = (2 / sigma^3) (d_i^2 a + f G_{i+1} (2 d_i + 1))
// (exact at the truncation boundary too: a node outside [lo, hi] has G = 0 and both derivatives 0), so no derivative
// table is read and no global memory is touched inside the voxel loop.  Slot pairs run as packed FP32x2; the four
// sums of a slot stay in registers over the tile-frame, are reduced across the warp and added with fp64 atomics.
// A window wider than the staged capacity (extreme deformation) takes the table-reading loop (param_grad_generic).  Oracle: torch autograd over an extended
// restatement (oracle.ExtendedPort); every test of this file is labelled "extension, not reference parity".
namespace dnmf {

constexpr int kExtTX = 8, kExtTY = 8;

struct ExtParams {
  const float* resid;  // [B][X][Y][Z]
  const int* frame_ids;
  const float* beta;
  const float* C;
  const float2* tab[3];
  const float2* tabp[3];
  const float2* tabs[3];
  const int* rng;
  double* gpos;  // [K][3]
  double* gsig;  // [K]
  double scale;  // 2 / (B_global * N)
  int X, Y, Z, K, T;
  int tz, ntx, nty, ntz;
  const float* pos;    // [K][3]
  const float* sigma;  // [K]
  int wmax0, wmax1, wmax2;  // staged window capacity per axis (the fused kernel's: its tiles are at least as large)
  int fast_div;             // the exact 3-instruction division is verified for all three axes (verify_coord_kernel)
  float rcp0, rcp1, rcp2;
};

// Table-reading form (any window width): four slots at a time, the derivative tables from global memory.
__device__ __noinline__ void param_grad_generic(const ExtParams& p, const float* sBeta, const unsigned short* sList, int L,
                                                const float* __restrict__ resid, int t, int x0, int y0, int z0, int nz,
                                                int lane) {
  const int lx = lane & 7, lyb = lane >> 3;
  const int gx = x0 + lx;
  const float xf = (float)gx;
  const float sm1x = (float)(p.X - 1), sm1y = (float)(p.Y - 1), sm1z = (float)(p.Z - 1);
  const int sX3 = p.X + 3, sY3 = p.Y + 3, sZ3 = p.Z + 3;

  for (int j0 = 0; j0 < L; j0 += 4) {
    float acc[4][4];
    int kk[4];
    float ck[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      kk[s] = sList[min(j0 + s, L - 1)];
      ck[s] = (j0 + s < L) ? __ldg(p.C + (size_t)kk[s] * p.T + t) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[s][q] = 0.f;
    }
    for (int h = 0; h < kExtTY / kWarpY; ++h) {
      const int gy = y0 + h * kWarpY + lyb;
      const bool valid = (gx < p.X) && (gy < p.Y);
      const float yf = (float)gy;
      float c0[3], c1[3], c2[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float v = sBeta[d];
        v = fmaf(sBeta[3 + d], xf, v);
        v = fmaf(sBeta[6 + d], yf, v);
        v = fmaf(sBeta[12 + d], xf * xf, v);
        v = fmaf(sBeta[15 + d], yf * yf, v);
        v = fmaf(sBeta[21 + d], xf * yf, v);
        c0[d] = v;
        c1[d] = fmaf(sBeta[27 + d], yf, fmaf(sBeta[24 + d], xf, sBeta[9 + d]));
        c2[d] = sBeta[18 + d];
      }
      const float* rcol = resid + ((size_t)min(gx, p.X - 1) * p.Y + min(gy, p.Y - 1)) * p.Z + z0;
      for (int zz = 0; zz < nz; ++zz) {
        const float zf = (float)(z0 + zz);
        int i0, i1, i2;
        float f0, f1, f2;
        split_coord(sample_coord(fmaf(zf, fmaf(zf, c2[0], c1[0]), c0[0]), sm1x), p.X, i0, f0);
        split_coord(sample_coord(fmaf(zf, fmaf(zf, c2[1], c1[1]), c0[1]), sm1y), p.Y, i1, f1);
        split_coord(sample_coord(fmaf(zf, fmaf(zf, c2[2], c1[2]), c0[2]), sm1z), p.Z, i2, f2);
        const float r = valid ? __ldg(rcol + zz) : 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const size_t k = (size_t)kk[s];
          const float2 e0 = __ldg(p.tab[0] + k * sX3 + (i0 + 2)), e1 = __ldg(p.tab[1] + k * sY3 + (i1 + 2)),
                       e2 = __ldg(p.tab[2] + k * sZ3 + (i2 + 2));
          const float2 q0 = __ldg(p.tabp[0] + k * sX3 + (i0 + 2)), q1 = __ldg(p.tabp[1] + k * sY3 + (i1 + 2)),
                       q2 = __ldg(p.tabp[2] + k * sZ3 + (i2 + 2));
          const float2 s0 = __ldg(p.tabs[0] + k * sX3 + (i0 + 2)), s1 = __ldg(p.tabs[1] + k * sY3 + (i1 + 2)),
                       s2 = __ldg(p.tabs[2] + k * sZ3 + (i2 + 2));
          const float a0 = fmaf(f0, e0.y, e0.x), a1 = fmaf(f1, e1.y, e1.x), a2 = fmaf(f2, e2.y, e2.x);
          const float p0 = fmaf(f0, q0.y, q0.x), p1 = fmaf(f1, q1.y, q1.x), p2 = fmaf(f2, q2.y, q2.x);
          const float g0 = fmaf(f0, s0.y, s0.x), g1 = fmaf(f1, s1.y, s1.x), g2 = fmaf(f2, s2.y, s2.x);
          const float w = r * ck[s];
          const float a12 = a1 * a2, a02 = a0 * a2, a01 = a0 * a1;
          acc[s][0] = fmaf(w, p0 * a12, acc[s][0]);
          acc[s][1] = fmaf(w, p1 * a02, acc[s][1]);
          acc[s][2] = fmaf(w, p2 * a01, acc[s][2]);
          acc[s][3] = fmaf(w, fmaf(g0, a12, fmaf(g1, a02, g2 * a01)), acc[s][3]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[s][q] = warp_sum(acc[s][q]);
      if (lane == 0 && j0 + s < L) {
        const int k = kk[s];
        atomicAdd(p.gpos + (size_t)k * 3 + 0, (double)acc[s][0] * p.scale);
        atomicAdd(p.gpos + (size_t)k * 3 + 1, (double)acc[s][1] * p.scale);
        atomicAdd(p.gpos + (size_t)k * 3 + 2, (double)acc[s][2] * p.scale);
        atomicAdd(p.gsig + k, (double)acc[s][3] * p.scale);
      }
    }
  }
}

constexpr int kExtGroup = 8;  // slots per pass over the tile (four packed pairs)

__global__ void __launch_bounds__(32, 16) param_grad_kernel(const __grid_constant__ ExtParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wsum = p.wmax0 + p.wmax1 + p.wmax2;
  float4* sTab = reinterpret_cast<float4*>(smem_raw);                         // [entry][4 pairs]: (G_2p, G_2p+1, D_2p, D_2p+1)
  float* sR = reinterpret_cast<float*>(sTab + (size_t)wsum * (kExtGroup / 2));  // residual tile [8][RS]
  const int RS = kExtTY * p.tz + 4;
  float* sBeta = sR + kExtTX * RS;                                            // 32 floats
  int* sInt = reinterpret_cast<int*>(sBeta + 32);                             // 8 ints
  unsigned short* sList = reinterpret_cast<unsigned short*>(sInt + 8);

  const int lane = threadIdx.x;
  const int bx = blockIdx.x, by = blockIdx.y;
  const int b = (int)blockIdx.z / p.ntz, bz = (int)blockIdx.z - b * p.ntz;
  const int t = p.frame_ids[b];
  const int x0 = bx * kExtTX, y0 = by * kExtTY, z0 = bz * p.tz;
  const int nx = min(kExtTX, p.X - x0), ny = min(kExtTY, p.Y - y0), nz = min(p.tz, p.Z - z0);
  const size_t N = (size_t)p.X * p.Y * p.Z;
  const float* __restrict__ resid = p.resid + (size_t)b * N;

  if (lane < 30) sBeta[lane] = p.beta[(size_t)lane * p.T + t];
  // residual tile -> shared memory first (rows of ny * Z contiguous floats when the tile spans the depth): its latency
  // overlaps the window and the list
  for (int lxx = 0; lxx < nx; ++lxx) {
    if (nz == p.Z) {
      const float* src = resid + ((size_t)(x0 + lxx) * p.Y + y0) * p.Z;
      for (int e = lane; e < ny * p.Z; e += 32) sR[lxx * RS + e] = __ldg(src + e);
    } else {
      for (int e = lane; e < ny * nz; e += 32) {
        const int ly = e / nz, zz = e - ly * nz;
        sR[lxx * RS + ly * nz + zz] = __ldg(resid + ((size_t)(x0 + lxx) * p.Y + (y0 + ly)) * p.Z + z0 + zz);
      }
    }
  }
  __syncwarp();
  if (lane < 3) {
    const int s = lane == 0 ? p.X : (lane == 1 ? p.Y : p.Z);
    int wlo, whi;
    tile_window_axis(sBeta + lane, 3, (float)x0, (float)y0, (float)z0, (float)(x0 + nx - 1), (float)(y0 + ny - 1),
                     (float)(z0 + nz - 1), s, wlo, whi);
    sInt[lane] = wlo;
    sInt[3 + lane] = whi;
  }
  __syncwarp();
  int wlo[3], whi[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    wlo[d] = sInt[d];
    whi[d] = sInt[3 + d];
  }
  int L = 0;
  for (int k0 = 0; k0 < p.K; k0 += 32) {
    const int k = k0 + lane;
    const bool ok = (k < p.K) && neuron_in_window(p.rng + (size_t)k * 6, wlo, whi);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) sList[L + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
    L += __popc(m);
  }
  __syncwarp();
  if (L == 0) return;
  const int W0 = whi[0] - wlo[0] + 1, W1 = whi[1] - wlo[1] + 1, W2 = whi[2] - wlo[2] + 1;
  if (W0 > p.wmax0 || W1 > p.wmax1 || W2 > p.wmax2) {
    param_grad_generic(p, sBeta, sList, L, resid, t, x0, y0, z0, nz, lane);
    return;
  }
  const int lx = lane & 7, lyb = lane >> 3;
  const int gx = x0 + lx;
  const float xf = (float)gx;
  const float sm1x = (float)(p.X - 1), sm1y = (float)(p.Y - 1), sm1z = (float)(p.Z - 1);
  const int sX3 = p.X + 3, sY3 = p.Y + 3, sZ3 = p.Z + 3;
  const int Wt = W0 + W1 + W2;
  constexpr int NPR = kExtGroup / 2;

  for (int j0 = 0; j0 < L; j0 += kExtGroup) {
    const int ng = min(kExtGroup, L - j0);
    __syncwarp();  // the previous group's voxel loop is done with the slices
    // ---- slices of slots j0 .. j0 + ng - 1: one lane per table entry, every pair of the group ----
    for (int e = lane; e < Wt; e += 32) {
      const float2* src;
      int row, ent;
      if (e < W0) {
        src = p.tab[0] + (wlo[0] + 2 + e);
        row = sX3;
        ent = e;
      } else if (e < W0 + W1) {
        src = p.tab[1] + (wlo[1] + 2 + (e - W0));
        row = sY3;
        ent = p.wmax0 + (e - W0);
      } else {
        src = p.tab[2] + (wlo[2] + 2 + (e - W0 - W1));
        row = sZ3;
        ent = p.wmax0 + p.wmax1 + (e - W0 - W1);
      }
      float2 va[NPR], vb[NPR];
#pragma unroll
      for (int pp = 0; pp < NPR; ++pp) {
        va[pp] = vb[pp] = make_float2(0.f, 0.f);
        if (2 * pp < ng) va[pp] = __ldg(src + (size_t)sList[j0 + 2 * pp] * row);
        if (2 * pp + 1 < ng) vb[pp] = __ldg(src + (size_t)sList[j0 + 2 * pp + 1] * row);
      }
#pragma unroll
      for (int pp = 0; pp < NPR; ++pp) sTab[ent * NPR + pp] = make_float4(va[pp].x, vb[pp].x, va[pp].y, vb[pp].y);
    }
    // ---- per-pair constants: positions, traces (0 for the padding slot of an odd group) ----
    float2 pos[NPR][3], ck[NPR];
#pragma unroll
    for (int pp = 0; pp < NPR; ++pp) {
      const int ka = sList[min(j0 + 2 * pp, L - 1)], kb = sList[min(j0 + 2 * pp + 1, L - 1)];
#pragma unroll
      for (int d = 0; d < 3; ++d) pos[pp][d] = make_float2(__ldg(p.pos + (size_t)ka * 3 + d), __ldg(p.pos + (size_t)kb * 3 + d));
      ck[pp] = make_float2(2 * pp < ng ? __ldg(p.C + (size_t)ka * p.T + t) : 0.f,
                           2 * pp + 1 < ng ? __ldg(p.C + (size_t)kb * p.T + t) : 0.f);
    }
    float2 acc[NPR][4];
#pragma unroll
    for (int pp = 0; pp < NPR; ++pp)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[pp][q] = make_float2(0.f, 0.f);
    __syncwarp();
    const int npr = (ng + 1) >> 1;
    for (int h = 0; h < kExtTY / kWarpY; ++h) {
      const int ly = h * kWarpY + lyb, gy = y0 + ly;
      const bool valid = (gx < p.X) && (gy < p.Y);
      const float yf = (float)gy;
      float c0[3], c1[3], c2[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float v = sBeta[d];
        v = fmaf(sBeta[3 + d], xf, v);
        v = fmaf(sBeta[6 + d], yf, v);
        v = fmaf(sBeta[12 + d], xf * xf, v);
        v = fmaf(sBeta[15 + d], yf * yf, v);
        v = fmaf(sBeta[21 + d], xf * yf, v);
        c0[d] = v;
        c1[d] = fmaf(sBeta[27 + d], yf, fmaf(sBeta[24 + d], xf, sBeta[9 + d]));
        c2[d] = sBeta[18 + d];
        if (p.fast_div) {  // exact doubling: the z loop then evaluates 2q, what sample_coord_fast expects
          c0[d] += c0[d];
          c1[d] += c1[d];
          c2[d] += c2[d];
        }
      }
      const float* rcol = sR + lx * RS + ly * nz;
      for (int zz = 0; zz < nz; ++zz) {
        const float zf = (float)(z0 + zz);
        int i0, i1, i2;
        float f0, f1, f2;
        const float q0 = fmaf(zf, fmaf(zf, c2[0], c1[0]), c0[0]), q1 = fmaf(zf, fmaf(zf, c2[1], c1[1]), c0[1]),
                    q2 = fmaf(zf, fmaf(zf, c2[2], c1[2]), c0[2]);
        if (p.fast_div) {
          split_coord(sample_coord_fast(q0, sm1x, p.rcp0, 0.5f * sm1x), p.X, i0, f0);
          split_coord(sample_coord_fast(q1, sm1y, p.rcp1, 0.5f * sm1y), p.Y, i1, f1);
          split_coord(sample_coord_fast(q2, sm1z, p.rcp2, 0.5f * sm1z), p.Z, i2, f2);
        } else {
          split_coord(sample_coord(q0, sm1x), p.X, i0, f0);
          split_coord(sample_coord(q1, sm1y), p.Y, i1, f1);
          split_coord(sample_coord(q2, sm1z), p.Z, i2, f2);
        }
        const float r = valid ? rcol[zz] : 0.f;
        // the conservative window covers every sample of the tile; the clamp only bounds the lanes outside the volume
        const int o0 = min(max(i0 - wlo[0], 0), W0 - 1), o1 = min(max(i1 - wlo[1], 0), W1 - 1),
                  o2 = min(max(i2 - wlo[2], 0), W2 - 1);
        const float n0 = (float)(o0 + wlo[0]), n1 = (float)(o1 + wlo[1]), n2 = (float)(o2 + wlo[2]);
        const float4* tx = sTab + (size_t)o0 * NPR;
        const float4* ty = sTab + (size_t)(p.wmax0 + o1) * NPR;
        const float4* tz = sTab + (size_t)(p.wmax0 + p.wmax1 + o2) * NPR;
        const float2 ff[3] = {make_float2(f0, f0), make_float2(f1, f1), make_float2(f2, f2)};
        const float2 nn[3] = {make_float2(n0, n0), make_float2(n1, n1), make_float2(n2, n2)};
        const float2 rr = make_float2(r, r);
#pragma unroll
        for (int pp = 0; pp < NPR; ++pp) {
          if (pp < npr) {  // warp-uniform
            const float4 e[3] = {tx[pp], ty[pp], tz[pp]};
            float2 a[3], pn[3], sn[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
              a[d] = __ffma2_rn(ff[d], make_float2(e[d].z, e[d].w), make_float2(e[d].x, e[d].y));  // footprint factor
            // both neurons of the pair vanish on all 32 voxels of this plane (cut-off): nothing to add
            {
              const float2 a012 = __fmul2_rn(__fmul2_rn(a[0], a[1]), a[2]);
              if (!__any_sync(0xffffffffu, (a012.x != 0.f || a012.y != 0.f) && r != 0.f)) continue;
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float2 G = make_float2(e[d].x, e[d].y), D = make_float2(e[d].z, e[d].w);
              const float2 fg = __fmul2_rn(ff[d], __fadd2_rn(G, D));            // f G_{i+1}
              const float2 dd = __fadd2_rn(nn[d], make_float2(-pos[pp][d].x, -pos[pp][d].y));  // d_i = i - pos
              pn[d] = __ffma2_rn(dd, a[d], fg);                                 // (sigma^2 / 2) lerp(dG/dpos)
              // (sigma^3 / 2) lerp(dG/dsigma) = d^2 a + f G_{i+1} (2 d + 1) = d (d a + f G_{i+1}) + f G_{i+1} (d + 1)
              sn[d] = __ffma2_rn(dd, pn[d], __fmul2_rn(fg, __fadd2_rn(dd, make_float2(1.f, 1.f))));
            }
            const float2 a12 = __fmul2_rn(a[1], a[2]), a02 = __fmul2_rn(a[0], a[2]), a01 = __fmul2_rn(a[0], a[1]);
            const float2 w = __fmul2_rn(rr, ck[pp]);
            acc[pp][0] = __ffma2_rn(__fmul2_rn(w, pn[0]), a12, acc[pp][0]);
            acc[pp][1] = __ffma2_rn(__fmul2_rn(w, pn[1]), a02, acc[pp][1]);
            acc[pp][2] = __ffma2_rn(__fmul2_rn(w, pn[2]), a01, acc[pp][2]);
            float2 sg = __fmul2_rn(sn[0], a12);
            sg = __ffma2_rn(sn[1], a02, sg);
            sg = __ffma2_rn(sn[2], a01, sg);
            acc[pp][3] = __ffma2_rn(w, sg, acc[pp][3]);
          }
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < NPR; ++pp) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[pp][q].x = warp_sum(acc[pp][q].x);
        acc[pp][q].y = warp_sum(acc[pp][q].y);
      }
      if (lane == 0) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (2 * pp + hh < ng) {
            const int k = sList[j0 + 2 * pp + hh];
            const double sg = (double)__ldg(p.sigma + k);
            const double cp = 2.0 / (sg * sg) * p.scale, cs = 2.0 / (sg * sg * sg) * p.scale;
            atomicAdd(p.gpos + (size_t)k * 3 + 0, (double)(hh ? acc[pp][0].y : acc[pp][0].x) * cp);
            atomicAdd(p.gpos + (size_t)k * 3 + 1, (double)(hh ? acc[pp][1].y : acc[pp][1].x) * cp);
            atomicAdd(p.gpos + (size_t)k * 3 + 2, (double)(hh ? acc[pp][2].y : acc[pp][2].x) * cp);
            atomicAdd(p.gsig + k, (double)(hh ? acc[pp][3].y : acc[pp][3].x) * cs);
          }
        }
      }
    }
  }
}

__global__ void ext_finish_kernel(const double* __restrict__ sumr, int B, double scale, double* __restrict__ gbg) {
  if (threadIdx.x < 32) {
    double acc = 0.0;
    for (int j = threadIdx.x; j < B; j += 32) acc += sumr[j];
    acc = warp_sum_d(acc);
    if (threadIdx.x == 0) *gbg = acc * scale;
  }
}

}  // namespace dnmf

using namespace dnmf;

extern "C" int dnmf_ext_enable(dnmf_ctx* c) {
  if (!c) return fail("dnmf_ext_enable: ctx is NULL");
  CU(cudaSetDevice(c->device));
  const int s[3] = {c->X, c->Y, c->Z};
  for (int d = 0; d < 3; ++d) {
    if (!c->d_tab_dpos[d]) CU(cudaMalloc((void**)&c->d_tab_dpos[d], (size_t)c->K * (s[d] + 3) * sizeof(float2)));
    if (!c->d_tab_dsig[d]) CU(cudaMalloc((void**)&c->d_tab_dsig[d], (size_t)c->K * (s[d] + 3) * sizeof(float2)));
  }
  c->have_footprints = false;  // tables must be rebuilt so that the derivative tables exist
  return 0;
}

// background_dev != NULL: the scalar background is read from the device (the context's own, dnmf_ext_step_begin)
static int ext_loss_grad_impl(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B, int B_global,
                              const float* beta_dev, const float* C_dev, float background, const float* background_dev,
                              float* grad_beta_dev, double* sse_dev, double* gpos_dev, double* gsig_dev,
                              double* gbg_dev, cudaStream_t st) {
  if (!c->d_tab_dpos[0]) return fail("dnmf_ext_loss_grad: call dnmf_ext_enable (then dnmf_set_footprints) first");
  if (B < 1 || B_global < B) return fail("dnmf_ext_loss_grad: need 1 <= B <= B_global");
  CU(cudaSetDevice(c->device));
  if (check_sticky(c, "dnmf_ext_loss_grad") || sanitize_ids(c, frame_ids_dev, B, st, &frame_ids_dev)) return 1;
  FitParams p;
  if (fill_fit_params(c, p, frames_dev, frame_ids_dev, B, beta_dev, C_dev, st)) return 1;
  if (ensure(&c->d_resid, &c->resid_cap, (size_t)B * c->N)) return 1;
  if (ensure(&c->d_sumr, &c->sumr_cap, (size_t)B)) return 1;
  p.yhat = c->d_resid;
  p.bg = background;
  p.bg_dev = background_dev;
  if (dispatch_fit<2>(c, p, B, st)) return 1;
  const int nt = c->ntx * c->nty * c->ntz;
  const double scale = 2.0 / ((double)B_global * (double)c->N);
  reduce_partials_kernel<<<B, 256, 0, st>>>(c->d_partials, frame_ids_dev, B, nt, c->T, scale, grad_beta_dev, sse_dev,
                                            c->d_sumr, nullptr, c->d_id_flags);
  CU(cudaGetLastError());
  ext_finish_kernel<<<1, 32, 0, st>>>(c->d_sumr, B, scale, gbg_dev);
  CU(cudaGetLastError());
  CU(cudaMemsetAsync(gpos_dev, 0, (size_t)c->K * 3 * sizeof(double), st));
  CU(cudaMemsetAsync(gsig_dev, 0, (size_t)c->K * sizeof(double), st));
  ExtParams e;
  e.resid = c->d_resid;
  e.frame_ids = frame_ids_dev;
  e.beta = beta_dev;
  e.C = C_dev;
  for (int d = 0; d < 3; ++d) {
    e.tab[d] = c->d_tab[d];
    e.tabp[d] = c->d_tab_dpos[d];
    e.tabs[d] = c->d_tab_dsig[d];
  }
  e.rng = c->d_rng;
  e.pos = c->d_pos;
  e.sigma = c->d_sigma;
  e.wmax0 = c->wmax[0];
  e.wmax1 = c->wmax[1];
  e.wmax2 = c->wmax[2];
  e.fast_div = c->fast_div;
  e.rcp0 = c->rcp[0];
  e.rcp1 = c->rcp[1];
  e.rcp2 = c->rcp[2];
  e.gpos = gpos_dev;
  e.gsig = gsig_dev;
  e.scale = scale;
  e.X = c->X;
  e.Y = c->Y;
  e.Z = c->Z;
  e.K = c->K;
  e.T = c->T;
  e.tz = c->tz;
  e.ntx = (c->X + kExtTX - 1) / kExtTX;
  e.nty = (c->Y + kExtTY - 1) / kExtTY;
  e.ntz = c->ntz;
  if (e.nty > 65535) return fail("dnmf_ext_loss_grad: more than 65535 tiles along y");
  const size_t smem = (size_t)(c->wmax[0] + c->wmax[1] + c->wmax[2]) * (kExtGroup / 2) * 16 +
                      (size_t)kExtTX * (kExtTY * c->tz + 4) * 4 + 32 * 4 + 8 * 4 + (size_t)((c->K + 7) & ~7) * 2;
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(param_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int maxB = std::max(1, 65535 / e.ntz);
  for (int b0 = 0; b0 < B; b0 += maxB) {
    const int nb = std::min(maxB, B - b0);
    ExtParams q = e;
    q.frame_ids = frame_ids_dev + b0;
    q.resid = c->d_resid + (size_t)b0 * c->N;
    dim3 grid((unsigned)e.ntx, (unsigned)e.nty, (unsigned)(nb * e.ntz));
    param_grad_kernel<<<grid, 32, smem, st>>>(q);
    CU(cudaGetLastError());
  }
  c->counters[0] += 1;
  c->counters[1] += 1;
  return 0;
}

extern "C" int dnmf_ext_loss_grad(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                                  int B_global, const float* beta_dev, const float* C_dev, float background,
                                  float* grad_beta_dev, double* sse_dev, double* gpos_dev, double* gsig_dev,
                                  double* gbg_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev || !C_dev || !grad_beta_dev || !sse_dev || !gpos_dev || !gsig_dev || !gbg_dev)
    return fail("dnmf_ext_loss_grad: NULL argument");
  return ext_loss_grad_impl(c, frames_dev, frame_ids_dev, B, B_global, beta_dev, C_dev, background, nullptr,
                            grad_beta_dev, sse_dev, gpos_dev, gsig_dev, gbg_dev, (cudaStream_t)stream);
}

// ---- device-resident iteration of the extension: no host round trip between the gradient kernels, the caller's
// all-reduce of ONE packed buffer, the Adam steps and the table rebuild ------------------------------------------
namespace dnmf {

// packed[0 .. 3K) = dL/dpos, [3K .. 4K) = dL/dsigma, [4K] = dL/db (already in place), [4K+1] = sum of the batch SSE
__global__ void ext_pack_sse_kernel(const double* __restrict__ sse, int B, double* __restrict__ packed_sse) {
  double acc = 0.0;
  for (int j = threadIdx.x; j < B; j += 32) acc += sse[j];
  acc = warp_sum_d(acc);
  if (threadIdx.x == 0) *packed_sse = acc;
}

// Adam (torch.optim._single_tensor_adam's formula, fp32 state) on pos[K][3], sigma[K] and the background, gradients
// from the packed (all-reduced) buffer; sigma is kept >= sigma_min.
__global__ void ext_adam_kernel(float* __restrict__ pos, float* __restrict__ sigma, float* __restrict__ bg,
                                float* __restrict__ m, float* __restrict__ v, const double* __restrict__ packed, int K,
                                float w1, float b2, float w2, float bc2_sqrt, float eps, float ss_pos,
                                float ss_sigma, float ss_bg, float sigma_min) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > 4 * K) return;
  float* prm = i < 3 * K ? pos + i : (i < 4 * K ? sigma + (i - 3 * K) : bg);
  const float ss = i < 3 * K ? ss_pos : (i < 4 * K ? ss_sigma : ss_bg);
  const float g = (float)packed[i];
  float x = *prm, mi = m[i], vi = v[i];
  adam_update(x, mi, vi, g, w1, b2, w2, ss, bc2_sqrt, eps);   // the arithmetic of adam_kernel
  m[i] = mi;
  v[i] = vi;
  if (i >= 3 * K && i < 4 * K) x = fmaxf(x, sigma_min);
  *prm = x;
}

// The refreshed candidate lists must fit the id buffer they were first sized for (+50 %): if they do not, the offsets
// are emptied (no kernel reads past the buffer) and the context's sticky error is raised -- the next call fails loudly
// and dnmf_set_footprints re-sizes everything.
__global__ void ext_guard_candidates_kernel(long long* __restrict__ cand_off, int nt, long long cap, int* sticky) {
  const long long total = cand_off[nt];
  if (total <= cap) return;
  for (int i = threadIdx.x; i <= nt; i += blockDim.x) cand_off[i] = 0;
  if (threadIdx.x == 0) atomicOr(sticky, 8);
}

// Ranges, tables and the static candidate lists from the device-resident pos / sigma at the CURRENT launch geometry
// (no layout search, no synchronisation).
static int ext_rebuild_tables(dnmf_ctx* c, cudaStream_t st) {
  build_ranges_kernel<<<(c->K + 127) / 128, 128, 0, st>>>(c->d_pos, c->d_sigma, c->K, c->X, c->Y, c->Z, c->cutoff,
                                                          c->d_rng);
  CU(cudaGetLastError());
  const int s[3] = {c->X, c->Y, c->Z};
  for (int d = 0; d < 3; ++d) {
    dim3 grid((s[d] + 3 + 127) / 128, c->K);
    build_tables_kernel<<<grid, 128, 0, st>>>(c->d_pos, c->d_sigma, c->d_rng, c->K, s[d], d, c->d_tab[d],
                                              c->d_tab_dpos[d], c->d_tab_dsig[d]);
    CU(cudaGetLastError());
  }
  const int nt = c->ntx * c->nty * c->ntz;
  if (run_bin_count(c, c->d_identity_beta, 1, c->d_ids_zero, 1, c->d_tmp_counts, nullptr, st, c->cand_expand)) return 1;
  scan_counts_kernel<<<1, 1024, 0, st>>>(c->d_tmp_counts, nt, c->d_cand_off, nullptr);
  CU(cudaGetLastError());
  ext_guard_candidates_kernel<<<1, 256, 0, st>>>(c->d_cand_off, nt, c->cand_ids_cap, c->d_sticky);
  CU(cudaGetLastError());
  Geom g1 = geom_of(c);
  g1.T = 1;
  const int wpb = 8;
  bin_tiles_kernel<true><<<(unsigned)((nt + wpb - 1) / wpb), wpb * 32, 0, st>>>(
      g1, c->d_identity_beta, c->d_ids_zero, 1, c->d_rng, c->d_tmp_counts, c->d_cand_off, nullptr, c->d_cand_ids,
      c->cand_ids_cap, c->cand_expand);
  CU(cudaGetLastError());
  c->mu_capM = 0;
  c->mu_nbr_built = false;
  c->gc_valid = false;
  c->mu_fused_need = 0;
  c->mu_fused_off = 0;
  c->counters[3]++;
  return 0;
}

static int ext_state(dnmf_ctx* c) {
  if (!c->d_tab_dpos[0] || !c->have_footprints)
    return fail("dnmf_ext_step: call dnmf_ext_enable and dnmf_set_footprints first");
  if (!c->d_bg) {
    CU(cudaMalloc((void**)&c->d_bg, sizeof(float)));
    CU(cudaMemset(c->d_bg, 0, sizeof(float)));
    CU(cudaMalloc((void**)&c->d_ext_m, ((size_t)4 * c->K + 1) * sizeof(float)));
    CU(cudaMalloc((void**)&c->d_ext_v, ((size_t)4 * c->K + 1) * sizeof(float)));
    CU(cudaMemset(c->d_ext_m, 0, ((size_t)4 * c->K + 1) * sizeof(float)));
    CU(cudaMemset(c->d_ext_v, 0, ((size_t)4 * c->K + 1) * sizeof(float)));
  }
  return 0;
}

}  // namespace dnmf

extern "C" int dnmf_ext_set_params(dnmf_ctx* c, float background, int reset_adam_state, void* stream) {
  if (!c) return fail("dnmf_ext_set_params: ctx is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (ext_state(c)) return 1;
  CU(cudaMemcpyAsync(c->d_bg, &background, sizeof(float), cudaMemcpyHostToDevice, st));
  if (reset_adam_state) {
    CU(cudaMemsetAsync(c->d_ext_m, 0, ((size_t)4 * c->K + 1) * sizeof(float), st));
    CU(cudaMemsetAsync(c->d_ext_v, 0, ((size_t)4 * c->K + 1) * sizeof(float), st));
  }
  CU(cudaStreamSynchronize(st));  // `background` is a stack value
  return 0;
}

extern "C" int dnmf_ext_get_params(dnmf_ctx* c, float* pos_dev_out, float* sigma_dev_out, float* background_dev_out,
                                   void* stream) {
  if (!c) return fail("dnmf_ext_get_params: ctx is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (ext_state(c)) return 1;
  if (pos_dev_out) CU(cudaMemcpyAsync(pos_dev_out, c->d_pos, (size_t)c->K * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (sigma_dev_out) CU(cudaMemcpyAsync(sigma_dev_out, c->d_sigma, (size_t)c->K * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (background_dev_out) CU(cudaMemcpyAsync(background_dev_out, c->d_bg, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int dnmf_ext_step_begin(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                                   int B_global, const float* beta_dev, const float* C_dev, double* packed_dev,
                                   void* stream) {
  if (!c || !frame_ids_dev || !beta_dev || !C_dev || !packed_dev) return fail("dnmf_ext_step_begin: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (ext_state(c)) return 1;
  if (ensure(&c->d_sse, &c->sse_cap, (size_t)B)) return 1;
  const int K = c->K;
  if (ext_loss_grad_impl(c, frames_dev, frame_ids_dev, B, B_global, beta_dev, C_dev, 0.f, c->d_bg, c->d_grad, c->d_sse,
                         packed_dev, packed_dev + 3 * (size_t)K, packed_dev + 4 * (size_t)K, st))
    return 1;
  ext_pack_sse_kernel<<<1, 32, 0, st>>>(c->d_sse, B, packed_dev + 4 * (size_t)K + 1);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_ext_step_end(dnmf_ctx* c, float* beta_dev, float* m_dev, float* v_dev, double lr, double beta1,
                                 double beta2, double eps, int64_t step, int affine, const double* packed_dev,
                                 int B_global, double lr_pos, double lr_sigma, double lr_background, float sigma_min,
                                 double* loss_dev, void* stream) {
  if (!c || !beta_dev || !m_dev || !v_dev || !packed_dev) return fail("dnmf_ext_step_end: NULL argument");
  if (step < 1) return fail("dnmf_ext_step_end: step is 1-based");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (ext_state(c)) return 1;
  const int K = c->K;
  // Adam on beta with the gradient dnmf_ext_step_begin left in the context; the loss is the all-reduced SSE sum
  if (dnmf_adam_step(c, beta_dev, c->d_grad, m_dev, v_dev, lr, beta1, beta2, eps, step, affine, packed_dev + 4 * (size_t)K + 1,
                     1, B_global, loss_dev, stream))
    return 1;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  ext_adam_kernel<<<(4 * K + 1 + 127) / 128, 128, 0, st>>>(
      c->d_pos, c->d_sigma, c->d_bg, c->d_ext_m, c->d_ext_v, packed_dev, K, (float)(1.0 - beta1),
      (float)beta2, (float)(1.0 - beta2), (float)sqrt(bc2), (float)eps, (float)(lr_pos / bc1), (float)(lr_sigma / bc1),
      (float)(lr_background / bc1), sigma_min);
  CU(cudaGetLastError());
  return ext_rebuild_tables(c, st);
}

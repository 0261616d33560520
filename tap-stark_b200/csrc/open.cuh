// open.cuh -- the reduced-opening pass of TwoAdicFriPcs::open (fri/src/two_adic_pcs.rs:260-419), SURVEY row f1:
// the step between the commitments and FRI.  It re-reads the committed LDE, so doing it on the device removes the
// only reason to download the LDE.
//
//   inv_denoms        :677-720   1/(x_X - z) for the coset g*K_h in bit-reversed order (x base field, z extension)
//   interpolate_coset :358-369   ys = p(z) from the low coset ([MEM] p3-interpolation; here the barycentric form
//                                p(z) = ((z/g)^n - 1)/n * sum_i e_i x_i / (z - x_i), exact in the field)
//   reduce rows       :371-381   ro[X] += alpha^off * (sum_i alpha^i p_i[X] - sum_i alpha^i y_i) * inv_denom[X]
#pragma once
#include "field.cuh"

namespace opn {

struct RootPows {
    uint32_t v[28];  // v[k] = w_h^(2^k), Montgomery
};
TS_D uint32_t pow_from_table(const RootPows &rp, uint32_t e) {
    uint32_t acc = bb::MONTY_ONE;
    for (int k = 0; e; k++, e >>= 1)
        if (e & 1) acc = bb::mmul(acc, rp.v[k]);
    return acc;
}
TS_D uint32_t brev_bits(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

TS_D uint32_t bb_inv(uint32_t a) {  // a^(p-2), Montgomery in/out; p - 2 = 0x77ffffff
    uint32_t r = bb::MONTY_ONE, b = a;
    uint32_t e = bb::P - 2;
    while (e) {
        if (e & 1) r = bb::mmul(r, b);
        b = bb::mmul(b, b);
        e >>= 1;
    }
    return r;
}
// inverse in F_p[x]/(x^4 - W) through the tower F_p[y]/(y^2 - W), y = x^2 (one base-field inversion)
TS_D ef::E4 ef_inv(const ef::E4 &a) {
    using bb::add;
    using bb::mmul;
    using bb::sub;
    const uint32_t W = bb::MONTY_W;
    const uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    const uint32_t a1a3 = mmul(a1, a3), a0a2 = mmul(a0, a2);
    // norm to F_p[y]: n0 + n1 y = A^2 - y B^2, A = a0 + a2 y, B = a1 + a3 y
    const uint32_t n0 = sub(add(mmul(a0, a0), mmul(W, mmul(a2, a2))), mmul(W, add(a1a3, a1a3)));
    const uint32_t n1 = sub(add(a0a2, a0a2), add(mmul(a1, a1), mmul(W, mmul(a3, a3))));
    const uint32_t d = bb_inv(sub(mmul(n0, n0), mmul(W, mmul(n1, n1))));
    const uint32_t m0 = mmul(n0, d), m1 = bb::neg(mmul(n1, d));
    ef::E4 r;
    r.c[0] = add(mmul(a0, m0), mmul(W, mmul(a2, m1)));
    r.c[2] = add(mmul(a0, m1), mmul(a2, m0));
    r.c[1] = bb::neg(add(mmul(a1, m0), mmul(W, mmul(a3, m1))));
    r.c[3] = bb::neg(add(mmul(a1, m1), mmul(a3, m0)));
    return r;
}
TS_D ef::E4 ef_mul_full(const ef::E4 &a, const ef::E4 &b) { return ef::mul(a, ef::prepare(b)); }

// out[X] = 1 / (g * w_h^bitrev(X) - z).
// Only the constant coefficient of the denominator depends on X, so the norm to F_p of a0 + a1 x + a2 x^2 + a3 x^3 (through
// F_p[y]/(y^2 - W): n0 + n1 y with n0 = a0^2 + C0, n1 = a0 C1 - C2) costs five multiplications per point; a thread takes the eight
// points X = j h/8 + t (bitrev(X) = 8 bitrev(t) + bitrev3(j): one table power per thread, coalesced stores across t) and inverts
// their norms with ONE field inversion (Montgomery's trick).  About 30 multiplications per point instead of 105.
__global__ void __launch_bounds__(256) inv_denoms_kernel(uint4 *out, int log_h, uint32_t g_monty, RootPows rp, ef::E4 z) {
    using bb::add;
    using bb::mmul;
    using bb::neg;
    using bb::sub;
    const size_t h = (size_t)1 << log_h;
    const uint32_t W = bb::MONTY_W;
    const uint32_t a1 = neg(z.c[1]), a2 = neg(z.c[2]), a3 = neg(z.c[3]);
    const uint32_t a1a3 = mmul(a1, a3);
    const uint32_t C0 = sub(mmul(W, mmul(a2, a2)), mmul(W, add(a1a3, a1a3)));
    const uint32_t C1 = add(a2, a2), C2 = add(mmul(a1, a1), mmul(W, mmul(a3, a3)));
    const uint32_t Wa2 = mmul(W, a2), Wa3 = mmul(W, a3);
    auto finish = [&](uint32_t a0, uint32_t n0, uint32_t n1, uint32_t d) {
        const uint32_t m0 = mmul(n0, d), m1 = neg(mmul(n1, d));
        return make_uint4(add(mmul(a0, m0), mmul(Wa2, m1)), neg(add(mmul(a1, m0), mmul(Wa3, m1))),
                          add(mmul(a0, m1), mmul(a2, m0)), neg(add(mmul(a1, m1), mmul(a3, m0))));
    };
    if (log_h < 3) {
        for (size_t X = (size_t)blockIdx.x * blockDim.x + threadIdx.x; X < h; X += (size_t)gridDim.x * blockDim.x) {
            const uint32_t a0 = sub(mmul(g_monty, pow_from_table(rp, brev_bits((uint32_t)X, log_h))), z.c[0]);
            const uint32_t n0 = add(mmul(a0, a0), C0), n1 = sub(mmul(a0, C1), C2);
            out[X] = finish(a0, n0, n1, bb_inv(sub(mmul(n0, n0), mmul(W, mmul(n1, n1)))));
        }
        return;
    }
    // w^bitrev3(j), j = 0..7
    uint32_t cw[8];
    cw[0] = bb::MONTY_ONE, cw[1] = rp.v[2], cw[2] = rp.v[1], cw[3] = mmul(rp.v[2], rp.v[1]), cw[4] = rp.v[0];
    cw[5] = mmul(cw[1], cw[4]), cw[6] = mmul(cw[2], cw[4]), cw[7] = mmul(cw[3], cw[4]);
    RootPows rp8;  // (w^8)^(2^k)
    TS_UNROLL
    for (int k = 0; k < 25; k++) rp8.v[k] = rp.v[k + 3];
    rp8.v[25] = rp8.v[26] = rp8.v[27] = bb::MONTY_ONE;
    const size_t per = h >> 3;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < per; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t base = mmul(g_monty, pow_from_table(rp8, brev_bits((uint32_t)t, log_h - 3)));
        uint32_t a0[8], n0[8], n1[8], nrm[8], pre[8];
        uint32_t run = bb::MONTY_ONE;
        TS_UNROLL
        for (int j = 0; j < 8; j++) {
            a0[j] = sub(mmul(base, cw[j]), z.c[0]);
            n0[j] = add(mmul(a0[j], a0[j]), C0);
            n1[j] = sub(mmul(a0[j], C1), C2);
            nrm[j] = sub(mmul(n0[j], n0[j]), mmul(W, mmul(n1[j], n1[j])));
            pre[j] = run;                              // product of the nonzero norms before j
            run = nrm[j] ? mmul(run, nrm[j]) : run;    // a zero norm (z on the domain) yields 0 like a^(p-2), without poisoning the batch
        }
        uint32_t inv = bb_inv(run);
        TS_UNROLL
        for (int j = 7; j >= 0; j--) {
            const uint32_t d = nrm[j] ? mmul(inv, pre[j]) : 0u;
            inv = nrm[j] ? mmul(inv, nrm[j]) : inv;
            out[(size_t)j * per + t] = finish(a0[j], n0[j], n1[j], d);
        }
    }
}

// Barycentric column sums over the low coset: partial[cta][c] = sum_{r in the CTA's row blocks} e[r][c] * k_r,
// k_r = x_r * inv_denom[r] (the sign and the ((z/g)^n - 1)/n factor are applied by the caller).
// The 256 threads of a CTA are (256 >> tpr_log) row lanes x (1 << tpr_log) columns -- one lane for a 200-column trace
// (coalesced 800-byte row reads), 64 lanes for a 4-column quotient chunk -- and a CTA walks the row blocks grid-stride with
// its sums in registers, so the number of partials is the grid size (a few per SM), not n / 128: the second kernel used to
// add 16384 partials per column in ONE serial loop per thread, which was most of the time of Pcs::open at the C4 shape.
constexpr int BARY_RB = 128;
// t < 2 p 2^32  ->  congruent mod p 2^32 and below it (2 p^2 + p 2^32 < 2 p 2^32 < 2^64: two products fit on top of a folded sum)
TS_D uint64_t fold64(uint64_t t) {
    const uint32_t hi = (uint32_t)(t >> 32);
    return ((uint64_t)bb::umin32(hi, hi - bb::P) << 32) | (uint32_t)t;
}
__global__ void __launch_bounds__(256) bary_partial_kernel(const uint32_t *__restrict__ m, size_t n, uint32_t width, int log_h,
                                                          uint32_t g_monty, RootPows rp, const uint4 *__restrict__ inv_denoms,
                                                          uint4 *__restrict__ partial, int tpr_log) {
    TS_DYN_SMEM(uint32_t, ks);  // BARY_RB x 4 weights, then 256 x 4 words for the lane reduction
    uint32_t *red = ks + BARY_RB * 4;
    const uint32_t tpr = 1u << tpr_log, lanes = 256u >> tpr_log;
    const uint32_t lane = threadIdx.x >> tpr_log, cl = threadIdx.x & (tpr - 1);
    const size_t n_blocks = (n + BARY_RB - 1) / BARY_RB;
    for (uint32_t c0 = 0; c0 < width; c0 += tpr) {
        const uint32_t c = c0 + cl;
        uint64_t acc[4] = {0, 0, 0, 0};  // running sums of products, kept below p 2^32 (fold64) and reduced once at the end
        for (size_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
            const size_t r0 = blk * BARY_RB;
            __syncthreads();
            for (int i = threadIdx.x; i < BARY_RB; i += blockDim.x) {
                uint4 k = make_uint4(0, 0, 0, 0);
                if (r0 + i < n) {
                    const uint32_t x = bb::mmul(g_monty, pow_from_table(rp, brev_bits((uint32_t)(r0 + i), log_h)));
                    const uint4 d = inv_denoms[r0 + i];
                    k = make_uint4(bb::mmul(x, d.x), bb::mmul(x, d.y), bb::mmul(x, d.z), bb::mmul(x, d.w));
                }
                ks[4 * i + 0] = k.x; ks[4 * i + 1] = k.y; ks[4 * i + 2] = k.z; ks[4 * i + 3] = k.w;
            }
            __syncthreads();
            if (c < width)
                for (uint32_t i0 = 2 * lane; i0 < (uint32_t)BARY_RB; i0 += 8 * lanes) {
                    // four row pairs per step: the eight loads are issued before the first product (the kernel is bound by
                    // loads in flight, not by arithmetic)
                    uint32_t v[8];
                    TS_UNROLL
                    for (int q = 0; q < 4; q++) {
                        const uint32_t i = i0 + 2 * lanes * q;
                        const bool in = i < (uint32_t)BARY_RB;
                        v[2 * q] = in && r0 + i < n ? m[(r0 + i) * width + c] : 0u;
                        v[2 * q + 1] = in && r0 + i + 1 < n ? m[(r0 + i + 1) * width + c] : 0u;
                    }
                    TS_UNROLL
                    for (int q = 0; q < 4; q++) {
                        const uint32_t i = i0 + 2 * lanes * q;
                        if (i < (uint32_t)BARY_RB) {
                            TS_UNROLL
                            for (int k = 0; k < 4; k++)
                                acc[k] = fold64(bb::madw(v[2 * q + 1], ks[4 * i + 4 + k], bb::madw(v[2 * q], ks[4 * i + k], acc[k])));
                        }
                    }
                }
        }
        uint32_t a[4];
        TS_UNROLL
        for (int k = 0; k < 4; k++) red[4 * threadIdx.x + k] = a[k] = bb::redc(acc[k]);
        __syncthreads();
        if (lane == 0 && c < width) {
            for (uint32_t l = 1; l < lanes; l++) {
                TS_UNROLL
                for (int k = 0; k < 4; k++) a[k] = bb::add(a[k], red[4 * ((l << tpr_log) + cl) + k]);
            }
            partial[(size_t)blockIdx.x * width + c] = make_uint4(a[0], a[1], a[2], a[3]);
        }
    }
}
// The same sums for widths that are multiples of 4 (every committed trace): a thread owns FOUR columns and reads them with one
// 16-byte load per row, eight rows per batch.  The scalar kernel above is bound by bytes in flight (4 B x 8 loads per thread:
// 1.2 TB/s on the 2^21 x 200 low coset, measured); this one keeps 128 B per thread in flight.
// Threads = (256 >> tq_log) row lanes x (1 << tq_log) column quads; 16 running 64-bit sums per thread.
constexpr int BARY4_RB = 256;
__global__ void __launch_bounds__(256, 2) bary_partial4_kernel(const uint4 *__restrict__ m, size_t n, uint32_t quads, int log_h,
                                                               uint32_t g_monty, RootPows rp, const uint4 *__restrict__ inv_denoms,
                                                               uint4 *__restrict__ partial, int tq_log) {
    TS_DYN_SMEM(uint32_t, sm);  // BARY4_RB x 4 weights, then 256 x 16 words for the lane reduction
    uint4 *ks = reinterpret_cast<uint4 *>(sm);
    uint32_t *red = sm + BARY4_RB * 4;
    const uint32_t tq = 1u << tq_log, lanes = 256u >> tq_log;
    const uint32_t lane = threadIdx.x >> tq_log, cq = threadIdx.x & (tq - 1);
    const size_t n_blocks = (n + BARY4_RB - 1) / BARY4_RB;
    for (uint32_t q0 = 0; q0 < quads; q0 += tq) {
        const uint32_t q = q0 + cq;
        uint64_t acc[4][4];
        TS_UNROLL
        for (int j = 0; j < 4; j++) {
            TS_UNROLL
            for (int k = 0; k < 4; k++) acc[j][k] = 0;
        }
        for (size_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
            const size_t r0 = blk * BARY4_RB;
            __syncthreads();
            {
                const uint32_t i = threadIdx.x;
                uint4 k = make_uint4(0, 0, 0, 0);
                if (r0 + i < n) {
                    const uint32_t x = bb::mmul(g_monty, pow_from_table(rp, brev_bits((uint32_t)(r0 + i), log_h)));
                    const uint4 d = inv_denoms[r0 + i];
                    k = make_uint4(bb::mmul(x, d.x), bb::mmul(x, d.y), bb::mmul(x, d.z), bb::mmul(x, d.w));
                }
                ks[i] = k;
            }
            __syncthreads();
            if (q < quads)
                for (uint32_t i0 = lane; i0 < (uint32_t)BARY4_RB; i0 += 8 * lanes) {
                    uint4 v[8];
                    TS_UNROLL
                    for (int s_ = 0; s_ < 8; s_++) {
                        const uint32_t i = i0 + (uint32_t)s_ * lanes;
                        v[s_] = (i < (uint32_t)BARY4_RB && r0 + i < n) ? m[(r0 + i) * quads + q] : make_uint4(0, 0, 0, 0);
                    }
                    TS_UNROLL
                    for (int s_ = 0; s_ < 8; s_ += 2) {
                        const uint32_t i = i0 + (uint32_t)s_ * lanes, i2 = i + lanes;
                        if (i < (uint32_t)BARY4_RB) {
                            const uint4 ka = ks[i], kb = i2 < (uint32_t)BARY4_RB ? ks[i2] : make_uint4(0, 0, 0, 0);
                            const uint32_t va[4] = {v[s_].x, v[s_].y, v[s_].z, v[s_].w};
                            const uint32_t vb[4] = {v[s_ + 1].x, v[s_ + 1].y, v[s_ + 1].z, v[s_ + 1].w};
                            TS_UNROLL
                            for (int j = 0; j < 4; j++) {
                                acc[j][0] = fold64(bb::madw(vb[j], kb.x, bb::madw(va[j], ka.x, acc[j][0])));
                                acc[j][1] = fold64(bb::madw(vb[j], kb.y, bb::madw(va[j], ka.y, acc[j][1])));
                                acc[j][2] = fold64(bb::madw(vb[j], kb.z, bb::madw(va[j], ka.z, acc[j][2])));
                                acc[j][3] = fold64(bb::madw(vb[j], kb.w, bb::madw(va[j], ka.w, acc[j][3])));
                            }
                        }
                    }
                }
        }
        uint32_t a[4][4];
        __syncthreads();  // `red` of the previous column group has been read
        TS_UNROLL
        for (int j = 0; j < 4; j++) {
            TS_UNROLL
            for (int k = 0; k < 4; k++) red[16 * threadIdx.x + 4 * j + k] = a[j][k] = bb::redc(acc[j][k]);
        }
        __syncthreads();
        if (lane == 0 && q < quads) {
            for (uint32_t l = 1; l < lanes; l++) {
                TS_UNROLL
                for (int j = 0; j < 4; j++) {
                    TS_UNROLL
                    for (int k = 0; k < 4; k++) a[j][k] = bb::add(a[j][k], red[16 * ((l << tq_log) + cq) + 4 * j + k]);
                }
            }
            TS_UNROLL
            for (int j = 0; j < 4; j++)
                partial[(size_t)blockIdx.x * (4 * quads) + 4 * q + j] = make_uint4(a[j][0], a[j][1], a[j][2], a[j][3]);
        }
    }
}
// ys[c] = sum_cta partial[cta][c]; 32 columns x 8 partial lanes per CTA
__global__ void __launch_bounds__(256) bary_final_kernel(const uint4 *__restrict__ partial, size_t n_partials, uint32_t width,
                                                        uint4 *__restrict__ ys) {
    TS_DYN_SMEM(uint32_t, red);  // 256 x 4
    const uint32_t c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
    uint32_t acc[4] = {0, 0, 0, 0};
    if (c < width)
        for (size_t b = ty; b < n_partials; b += 8) {
            const uint4 p = partial[b * width + c];
            acc[0] = bb::add(acc[0], p.x); acc[1] = bb::add(acc[1], p.y);
            acc[2] = bb::add(acc[2], p.z); acc[3] = bb::add(acc[3], p.w);
        }
    TS_UNROLL
    for (int k = 0; k < 4; k++) red[4 * threadIdx.x + k] = acc[k];
    __syncthreads();
    if (ty == 0 && c < width) {
        for (uint32_t l = 1; l < 8; l++) {
            TS_UNROLL
            for (int k = 0; k < 4; k++) acc[k] = bb::add(acc[k], red[4 * (32 * l + threadIdx.x) + k]);
        }
        ys[c] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    }
}

// ro[X] += apo * (dot[X] - rys) * inv_denoms[X]
__global__ void __launch_bounds__(256) reduce_rows_kernel(const uint4 *__restrict__ dot, const uint4 *__restrict__ inv_denoms,
                                                         ef::E4 apo, ef::E4 rys, size_t h, uint4 *__restrict__ ro) {
    const ef::E4Const ka = ef::prepare(apo);
    for (size_t X = (size_t)blockIdx.x * blockDim.x + threadIdx.x; X < h; X += (size_t)gridDim.x * blockDim.x) {
        const uint4 dv = dot[X], iv = inv_denoms[X], rv = ro[X];
        const ef::E4 d{{dv.x, dv.y, dv.z, dv.w}}, id{{iv.x, iv.y, iv.z, iv.w}}, r{{rv.x, rv.y, rv.z, rv.w}};
        const ef::E4 t = ef::mul(ef_mul_full(ef::sub(d, rys), id), ka);
        const ef::E4 o = ef::add(r, t);
        ro[X] = make_uint4(o.c[0], o.c[1], o.c[2], o.c[3]);
    }
}

// Query-phase gather: segment s copies n_words[s] words from src[s] to dst + dst_off[s].  One launch packs the opened rows
// and authentication paths of EVERY query of a tree into one buffer, which is read back with a single copy (round 1
// issued one 32-byte device-to-host copy per tree level and per row: hundreds to thousands per proof).
struct GatherSeg {
    const uint32_t *src;
    uint32_t dst_off, n_words;
};
__global__ void gather_segments_kernel(const GatherSeg *segs, uint32_t n_segs, uint32_t *dst) {
    for (uint32_t s = blockIdx.x; s < n_segs; s += gridDim.x) {
        const GatherSeg g = segs[s];
        for (uint32_t i = threadIdx.x; i < g.n_words; i += blockDim.x) dst[g.dst_off + i] = g.src[i];
    }
}

}  // namespace opn

// fri_tail.cuh -- the last rounds of bf_commit_phase (fri/src/prover.rs:93-141) in ONE launch.
//
// Below a few thousand elements a commit-phase round is pure latency: leaf hash + tree + fold are 3-4 launches of a few
// microseconds each, then a 32-byte root read-back, the host sponge and the next launch (~43 us per round on a B200,
// profiles/r01/config_sweep_b200.jsonl).  One CTA does all of it here, the Fiat-Shamir sponge included: after a digest
// has been observed the BfChallenger's state is (digest, previous squeeze) (basic/src/challenger/mod.rs:151-194:
// duplexing writes the 8 observed words over state[0..8], Blake3Permutation hashes the 64-byte state, the 32-byte
// result becomes state[8..16] and the output buffer, popped from the back), so per round
//     h_r = Blake3(root_r || h_{r-1}),   beta_r = (h_r[7], h_r[6], h_r[5], h_r[4]) mod p
// -- the same 64-byte compression as a tree node.  The host replays the sponge from the returned roots (22 host hashes)
// so its challenger stays the source of truth for grinding and the query phase.
#pragma once
#include "fold.cuh"
#include "hash.cuh"

namespace ftail {

constexpr int MAX_ROUNDS = 16;
constexpr int NT = 1024;
constexpr int MAX_LOG_LEN = 12;  // layers of at most 4096 extension elements

struct Params {
    const uint32_t *layer0;          // current layer: 2^log_len0 x 4 words, Montgomery
    int log_len0, rounds;
    uint32_t *layers[MAX_ROUNDS];    // layers[r]: output of round r (2^(log_len0 - r - 1) x 4) = leaves of round r + 1
    uint32_t *digests[MAX_ROUNDS];   // digests[r]: all tree layers of round r, leaves first (2 h_r - 1 nodes x 8 words)
    uint32_t prev_h[8];              // sponge state[8..16] on entry
    const uint32_t *prev_h_dev;      // != nullptr: read it from device memory instead (left by sponge_step_kernel)
    uint32_t inv_gen[MAX_LOG_LEN + 2];  // two_adic_generator(b)^-1, Montgomery
    uint32_t *roots_out;             // rounds x 8 words
};

__global__ void __launch_bounds__(NT) fri_tail_kernel(Params p) {
    TS_DYN_SMEM(uint32_t, sm);  // 12 words: sponge squeeze, beta / 2
    uint32_t *s_h = sm, *s_beta_half = sm + 8;
    const int tid = threadIdx.x;
    if (tid < 8) s_h[tid] = p.prev_h_dev ? p.prev_h_dev[tid] : p.prev_h[tid];
    __syncthreads();
    const uint32_t *cur = p.layer0;
    for (int r = 0; r < p.rounds; r++) {
        const int log_h = p.log_len0 - r - 1;
        const uint32_t h = 1u << log_h;
        uint32_t *dg = p.digests[r];
        // leaves: row t = (cur[2t], cur[2t+1]) as 8 canonical words (prover.rs:112)
        for (uint32_t t = tid; t < h; t += NT) {
            const uint4 a = reinterpret_cast<const uint4 *>(cur)[2 * t], b = reinterpret_cast<const uint4 *>(cur)[2 * t + 1];
            uint32_t m[16] = {bb::from_monty(a.x), bb::from_monty(a.y), bb::from_monty(a.z), bb::from_monty(a.w),
                              bb::from_monty(b.x), bb::from_monty(b.y), bb::from_monty(b.z), bb::from_monty(b.w),
                              0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t cv[8];
            b3::iv(cv);
            b3::compress(cv, m, 0, 32, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT);
            TS_UNROLL
            for (int i = 0; i < 8; i++) dg[(size_t)t * 8 + i] = cv[i];
        }
        __syncthreads();
        // tree: level of w nodes at node offset `off`, parents behind it
        uint32_t off = 0;
        for (uint32_t w = h; w > 1; w >>= 1) {
            for (uint32_t t = tid; t < w / 2; t += NT) {
                uint32_t l[8], rr[8], o[8];
                TS_UNROLL
                for (int i = 0; i < 8; i++) {
                    l[i] = dg[(size_t)(off + 2 * t) * 8 + i];
                    rr[i] = dg[(size_t)(off + 2 * t + 1) * 8 + i];
                }
                b3::compress_pair(l, rr, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT, o);
                TS_UNROLL
                for (int i = 0; i < 8; i++) dg[(size_t)(off + w + t) * 8 + i] = o[i];
            }
            off += w;
            __syncthreads();
        }
        // root -> sponge -> beta
        if (tid == 0) {
            uint32_t root[8], hp[8], hn[8];
            TS_UNROLL
            for (int i = 0; i < 8; i++) {
                root[i] = dg[(size_t)off * 8 + i];
                hp[i] = s_h[i];
                p.roots_out[r * 8 + i] = root[i];
            }
            b3::compress_pair(root, hp, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT, hn);
            TS_UNROLL
            for (int i = 0; i < 8; i++) s_h[i] = hn[i];
            TS_UNROLL
            for (int k = 0; k < 4; k++) {
                uint32_t v = hn[7 - k];  // pop from the back; u32 mod p: at most two subtractions
                v = v >= bb::P ? v - bb::P : v;
                v = v >= bb::P ? v - bb::P : v;
                s_beta_half[k] = bb::mmul(bb::to_monty(v), bb::MONTY_HALF);
            }
        }
        __syncthreads();
        // fold (two_adic_pcs.rs:116-147): out[t] = (lo + hi)/2 + (beta/2) * g_inv^bitrev(t) * (lo - hi)
        const ef::E4 half_beta{{s_beta_half[0], s_beta_half[1], s_beta_half[2], s_beta_half[3]}};
        const ef::E4Const hb = ef::prepare(half_beta);
        uint32_t *out = p.layers[r];
        for (uint32_t t = tid; t < h; t += NT) {
            uint32_t s = bb::MONTY_ONE;
            const uint32_t e = fold::brev_bits(t, log_h);
            for (int k = 0; k < log_h; k++)
                if ((e >> k) & 1u) s = bb::mmul(s, p.inv_gen[log_h + 1 - k]);  // g_inv^(2^k) = two_adic_generator(log_h+1-k)^-1
            const uint4 a = reinterpret_cast<const uint4 *>(cur)[2 * t], b = reinterpret_cast<const uint4 *>(cur)[2 * t + 1];
            const ef::E4 lo{{a.x, a.y, a.z, a.w}}, hi{{b.x, b.y, b.z, b.w}};
            const ef::E4 sum = ef::half(ef::add(lo, hi));
            const ef::E4 dif = ef::scale(ef::sub(lo, hi), s);
            const ef::E4 res = ef::add(sum, ef::mul(dif, hb));
            reinterpret_cast<uint4 *>(out)[t] = make_uint4(res.c[0], res.c[1], res.c[2], res.c[3]);
        }
        __syncthreads();
        cur = out;
    }
}

// One commit-phase round's Fiat-Shamir step on the device (same sponge identity as above): combines the n_sub sub-roots of
// a row-sharded layer (n_sub = 1: the root itself) into the layer root with the top log2(n_sub) tree levels, then
// h = Blake3(root || h_prev), beta = (h[7], h[6], h[5], h[4]) mod p.  Leaves the root for the host's replay, the new h for
// the next round and beta/2 (Montgomery) for this round's fold kernel -- nothing returns to the host between rounds.
__global__ void sponge_step_kernel(const uint32_t *sub_roots, int n_sub, uint32_t *h_state, uint32_t *root_out,
                                   uint32_t *half_beta_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t lvl[32][8];
    for (int i = 0; i < n_sub; i++)
        for (int k = 0; k < 8; k++) lvl[i][k] = sub_roots[i * 8 + k];
    for (int w = n_sub; w > 1; w >>= 1)
        for (int i = 0; i < w / 2; i++) {
            uint32_t o[8];
            b3::compress_pair(lvl[2 * i], lvl[2 * i + 1], b3::CHUNK_START | b3::CHUNK_END | b3::ROOT, o);
            for (int k = 0; k < 8; k++) lvl[i][k] = o[k];
        }
    uint32_t hp[8], hn[8];
    for (int k = 0; k < 8; k++) {
        hp[k] = h_state[k];
        root_out[k] = lvl[0][k];
    }
    b3::compress_pair(lvl[0], hp, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT, hn);
    for (int k = 0; k < 8; k++) h_state[k] = hn[k];
    for (int k = 0; k < 4; k++) {
        uint32_t v = hn[7 - k];
        v = v >= bb::P ? v - bb::P : v;
        v = v >= bb::P ? v - bb::P : v;
        half_beta_out[k] = bb::mmul(bb::to_monty(v), bb::MONTY_HALF);
    }
}

// ---- sub-root exchange through peer memory (row-sharded layers, one process per GPU) ------------------------------------
// Every rank owns a small MAILBOX that its peers map over CUDA IPC: roots[round][rank][8 words] and flags[round][rank].
// After building the subtree of round r a rank PUBLISHES its sub-root into all G mailboxes (NVLink stores, then a
// system-scope fence, then the flag = epoch); the sponge step of round r spins until all G flags of the round carry the
// epoch and then runs the top of the tree and the sponge exactly as sponge_step_kernel does.  No NCCL call and no host
// work separates the rounds: the collective is part of the kernels (the epoch, one per proving step, makes the mailbox
// reusable without clearing it).
constexpr int MAIL_ROUNDS = 32, MAIL_RANKS = 8;
constexpr int MAIL_FLAGS_OFF = MAIL_ROUNDS * MAIL_RANKS * 8;           // in words
constexpr int MAIL_WORDS = MAIL_FLAGS_OFF + MAIL_ROUNDS * MAIL_RANKS;  // 9216 words = 36 KiB
struct Mailboxes {
    uint32_t *box[MAIL_RANKS];
};
__global__ void publish_subroot_kernel(const uint32_t *sub_root, Mailboxes mb, int n_ranks, int my_rank, int round, uint32_t epoch) {
    const int d = threadIdx.x;
    if (d >= n_ranks || blockIdx.x != 0) return;
    volatile uint32_t *dst = mb.box[d] + ((size_t)round * MAIL_RANKS + my_rank) * 8;
    for (int k = 0; k < 8; k++) dst[k] = sub_root[k];
#ifndef TS_EMULATE
    __threadfence_system();
#endif
    volatile uint32_t *flag = mb.box[d] + MAIL_FLAGS_OFF + round * MAIL_RANKS + my_rank;
    *flag = epoch;
}
__global__ void sponge_step_mailbox_kernel(const uint32_t *mailbox, int n_ranks, int round, uint32_t epoch, uint32_t *h_state,
                                           uint32_t *root_out, uint32_t *half_beta_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const volatile uint32_t *flags = mailbox + MAIL_FLAGS_OFF + round * MAIL_RANKS;
    // bounded wait (about two seconds): a peer that never publishes must not hang the GPU; the sponge then runs on whatever
    // the mailbox holds and the transcript check of the caller fails loudly instead
    uint32_t spins = 0;
    for (int s = 0; s < n_ranks; s++) {
        while (flags[s] != epoch && spins < 10000000u) {
            spins++;
#if !defined(TS_EMULATE) && defined(__CUDA_ARCH__)
            __nanosleep(200);
#endif
        }
    }
#ifndef TS_EMULATE
    __threadfence_system();
#endif
    uint32_t lvl[MAIL_RANKS][8];
    const volatile uint32_t *roots = mailbox + (size_t)round * MAIL_RANKS * 8;
    for (int i = 0; i < n_ranks; i++)
        for (int k = 0; k < 8; k++) lvl[i][k] = roots[i * 8 + k];
    for (int w = n_ranks; w > 1; w >>= 1)
        for (int i = 0; i < w / 2; i++) {
            uint32_t o[8];
            b3::compress_pair(lvl[2 * i], lvl[2 * i + 1], b3::CHUNK_START | b3::CHUNK_END | b3::ROOT, o);
            for (int k = 0; k < 8; k++) lvl[i][k] = o[k];
        }
    uint32_t hp[8], hn[8];
    for (int k = 0; k < 8; k++) {
        hp[k] = h_state[k];
        root_out[k] = lvl[0][k];
    }
    b3::compress_pair(lvl[0], hp, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT, hn);
    for (int k = 0; k < 8; k++) h_state[k] = hn[k];
    for (int k = 0; k < 4; k++) {
        uint32_t v = hn[7 - k];
        v = v >= bb::P ? v - bb::P : v;
        v = v >= bb::P ? v - bb::P : v;
        half_beta_out[k] = bb::mmul(bb::to_monty(v), bb::MONTY_HALF);
    }
}

}  // namespace ftail

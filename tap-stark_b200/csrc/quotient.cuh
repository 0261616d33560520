// quotient.cuh -- quotient_values of uni_stark::prove (uni-stark/src/prover.rs:122-194), SURVEY row f3: the caller-side
// step between the trace commitment and the quotient commitment.  It reads the committed trace LDE where it lies
// (get_evaluations_on_domain, fri/src/two_adic_pcs.rs:247-258, is a VIEW of its first rows, so nothing is copied or
// downloaded) and writes the quotient chunks ready for the second ts_pcs_commit.
//
// The AIR arrives as a constraint program -- what `Air::eval` does to a ProverConstraintFolder
// (uni-stark/src/folder.rs:10-66), flattened to three-address code over base-field registers:
//
//     instruction = 4 words {op, dst, a, b};  operand = kind << 28 | index
//     kinds: REG r | LOCAL column | NEXT column | PUBLIC k | CONST k | SEL {0 is_first_row, 1 is_last_row, 2 is_transition}
//     ops:   ADD, SUB, MUL, NEG  (dst = a op b)        ASSERT_ZERO  (accumulator = accumulator * alpha + a, folder.rs:60-64)
//
// `when_first_row().assert_eq(x, y)` is ASSERT_ZERO(MUL(SEL 0, SUB(x, y))) exactly as p3's FilteredAirBuilder does it.
// Per quotient-domain point x_i = g w_m^i ([MEM] p3-commit TwoAdicMultiplicativeCoset::selectors_on_coset):
//     Z_H(x) = x^n - 1,  is_first_row = Z_H(x) / (x - 1),  is_last_row = Z_H(x) / (x - w_n^-1),
//     is_transition = x - w_n^-1,  quotient = accumulator / Z_H(x)                     (prover.rs:176-177)
// Z_H takes 2^(log m - log n) distinct values on the coset: two small host-built tables.
#pragma once
#include "field.cuh"
#include "open.cuh"

namespace quo {

constexpr uint32_t K_REG = 0, K_LOCAL = 1, K_NEXT = 2, K_PUBLIC = 3, K_CONST = 4, K_SEL = 5;
constexpr uint32_t OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_NEG = 3, OP_ASSERT_ZERO = 4;
constexpr int MAX_REGS = 64;
constexpr int MAX_ZH = 64;  // 2^(log m - log n) <= 2^6
constexpr int MAX_CHUNKS = 16;

struct Params {
    const uint32_t *lde;       // committed trace LDE: row t = evaluation at g w_N^brev(t); the first m rows are g H_m
    uint32_t width;
    int log_n, log_m;          // trace length, quotient-domain size
    int log_chunks;            // quotient_degree = 2^log_chunks chunks, chunk k = rows k, k + qd, ... (split_evals)
    const uint32_t *program;   // n_instr x 4
    uint32_t n_instr;
    const uint32_t *consts;    // Montgomery
    const uint32_t *publics;   // Montgomery
    ef::E4 alpha;
    opn::RootPows rp;          // w_m^(2^k)
    uint32_t g_monty;          // coset shift (generator)
    uint32_t wn_inv;           // w_n^-1
    uint32_t zh[MAX_ZH];       // Z_H(x_i) for i mod 2^(log m - log n)
    uint32_t zh_inv[MAX_ZH];
    uint32_t *out[MAX_CHUNKS]; // chunk k: (m >> log_chunks) x 4 words, natural order on its own coset
};

TS_D uint32_t operand(const Params &p, uint32_t code, const uint32_t *reg, size_t local_row, size_t next_row, const uint32_t *sel) {
    const uint32_t kind = code >> 28, idx = code & 0x0fffffffu;
    switch (kind) {
        case K_REG: return reg[idx];
        case K_LOCAL: return p.lde[local_row * p.width + idx];
        case K_NEXT: return p.lde[next_row * p.width + idx];
        case K_PUBLIC: return p.publics[idx];
        case K_CONST: return p.consts[idx];
        default: return sel[idx];
    }
}

// One thread per committed row t (natural quotient-domain index i = brev(t)), grid-stride.  The register file of the
// constraint program is per-thread local memory ON PURPOSE: a thread reads its two trace rows column by column, 32 threads of
// a warp touch 32 different rows per load, and only the L1 turns those 4-byte reads into one 128-byte line fetch per 32
// columns.  Moving the register file into shared memory (tried in round 2: reg[index][thread], conflict free) shrinks the
// L1 by the carve-out and made the kernel 3x SLOWER (width 60: 1.64 ms against 0.53 ms; profiles/r02/README.md).
__global__ void __launch_bounds__(128) quotient_values_kernel(Params p) {
    const size_t m = (size_t)1 << p.log_m;
    const uint32_t next_step = 1u << (p.log_m - p.log_n);                       // prover.rs:141-142
    const uint32_t zmask = next_step - 1;
    const uint32_t qd_mask = (1u << p.log_chunks) - 1;
    const ef::E4Const ka = ef::prepare(p.alpha);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = opn::brev_bits((uint32_t)t, p.log_m);
        const uint32_t i_next = (i + next_step) & (uint32_t)(m - 1);
        const size_t next_row = opn::brev_bits(i_next, p.log_m);
        // selectors at x = g w_m^i
        const uint32_t x = bb::mmul(p.g_monty, opn::pow_from_table(p.rp, i));
        const uint32_t zh = p.zh[i & zmask];
        const uint32_t d_first = bb::sub(x, bb::MONTY_ONE), d_last = bb::sub(x, p.wn_inv);
        const uint32_t inv_both = opn::bb_inv(bb::mmul(d_first, d_last));       // one inversion for both denominators
        uint32_t sel[3];
        sel[0] = bb::mmul(zh, bb::mmul(inv_both, d_last));
        sel[1] = bb::mmul(zh, bb::mmul(inv_both, d_first));
        sel[2] = d_last;
        uint32_t reg[MAX_REGS];
        ef::E4 acc{{0, 0, 0, 0}};
        for (uint32_t pc = 0; pc < p.n_instr; pc++) {
            const uint32_t op = p.program[4 * pc], dst = p.program[4 * pc + 1];
            const uint32_t a = operand(p, p.program[4 * pc + 2], reg, t, next_row, sel);
            if (op == OP_ASSERT_ZERO) {
                acc = ef::mul(acc, ka);
                acc.c[0] = bb::add(acc.c[0], a);
                continue;
            }
            if (op == OP_NEG) {
                reg[dst] = bb::neg(a);
                continue;
            }
            const uint32_t b = operand(p, p.program[4 * pc + 3], reg, t, next_row, sel);
            reg[dst] = op == OP_ADD ? bb::add(a, b) : op == OP_SUB ? bb::sub(a, b) : bb::mmul(a, b);
        }
        const uint32_t izh = p.zh_inv[i & zmask];
        uint32_t *o = p.out[i & qd_mask] + (size_t)(i >> p.log_chunks) * 4;
        *reinterpret_cast<uint4 *>(o) = make_uint4(bb::mmul(acc.c[0], izh), bb::mmul(acc.c[1], izh), bb::mmul(acc.c[2], izh),
                                                   bb::mmul(acc.c[3], izh));
    }
}

}  // namespace quo

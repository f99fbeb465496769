/* Gate descriptors shared by the C oracle prover.  TEST INFRASTRUCTURE. */
#ifndef ORACLE_GATES_H
#define ORACLE_GATES_H
#include "gl.h"
enum {
    G_NOOP = 0, G_CONSTANT, G_PUBLIC_INPUT, G_ARITHMETIC, G_ARITHMETIC_EXT, G_MUL_EXT, G_BASE_SUM, G_REDUCING,
    G_REDUCING_EXT, G_RANDOM_ACCESS, G_POSEIDON_MDS, G_POSEIDON, G_COSET_INTERPOLATION,
    G_U32_ARITHMETIC, G_U32_ADD_MANY, G_U32_SUBTRACTION, G_U32_RANGE_CHECK, G_U32_INTERLEAVE,
    G_UNINTERLEAVE_U32, G_UNINTERLEAVE_B32, G_COMPARISON, G_NUM_KINDS
};
typedef struct {
    uint32_t kind;
    uint32_t p[4];            /* parameters, meaning per kind (see gates.c) */
    uint32_t selector_index, group_start, group_end, row;   /* row = index of this gate in the sorted gate list */
    uint32_t num_constraints;
    const u64 *extra;         /* coset interpolation: barycentric weights */
} orc_gate;
int orc_eval_gate(const orc_gate *g, const u64 *consts, const u64 *wires, const u64 pi_hash[4], u64 *out);
void orc_eval_filtered(const orc_gate *gates, uint32_t num_gates, uint32_t num_selectors, uint32_t num_gate_constraints,
                       const u64 *consts, const u64 *wires, const u64 pi_hash[4], u64 *total, u64 *scratch);
#endif

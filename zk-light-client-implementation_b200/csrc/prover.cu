// b200_circuit_upload / b200_prove: placeholder until the full prover lands (returns B200_ERR_UNSUPPORTED).
#include "../../include/b200_plonky2.h"
#include "ctx.h"
extern "C" {
int b200_circuit_upload(b200_ctx *ctx, const b200_common *, const uint64_t *, const uint64_t[4], uint64_t *, b200_circuit **) {
    if (ctx) ctx->err = "b200_circuit_upload: not implemented yet"; return B200_ERR_UNSUPPORTED; }
void b200_circuit_free(b200_circuit *) {}
int b200_prove(b200_ctx *ctx, const b200_circuit *, const uint64_t *, const uint64_t *, size_t, uint64_t, uint8_t **, size_t *) {
    if (ctx) ctx->err = "b200_prove: not implemented yet"; return B200_ERR_UNSUPPORTED; }
int b200_prove_dev(b200_ctx *ctx, const b200_circuit *, const uint64_t *, const uint64_t *, size_t, uint64_t, uint8_t **, size_t *) {
    if (ctx) ctx->err = "b200_prove_dev: not implemented yet"; return B200_ERR_UNSUPPORTED; }
int b200_fri_fold(b200_ctx *ctx, const uint64_t *, uint32_t, uint32_t, const uint64_t[2], uint64_t, uint64_t *) {
    if (ctx) ctx->err = "b200_fri_fold: not implemented yet"; return B200_ERR_UNSUPPORTED; }
}

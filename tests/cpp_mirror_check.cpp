// Compiles and links the C++ host mirror against libb200plonky2.so; at run time checks the no-device error path
// (or, with a GPU, that a context can be created).  Built and run by tests/test_abi_exports.py.
#include "b200_plonky2.hpp"
#include <cstdio>
int main() {
    try {
        b200::Context ctx(0);
        std::printf("context ok\n");
        b200::CommonCircuitData cd;          // an inconsistent circuit must be rejected with an error, not a crash
        cd.c.num_wires = 4; cd.c.num_routed_wires = 8; cd.c.num_challenges = 2; cd.c.degree_bits = 3; cd.c.rate_bits = 3;
        cd.c.cap_height = 1; cd.c.proof_of_work_bits = 16; cd.c.quotient_degree_factor = 8; cd.c.num_partial_products = 0;
        uint64_t vals[8] = {0}, dg[4] = {0};
        try { b200::CircuitData bad(ctx, cd, vals, dg); std::printf("UNEXPECTED: accepted\n"); return 2; }
        catch (const b200::Error &e) { std::printf("rejected as expected: code %d\n", e.code); }
        return 0;
    } catch (const b200::Error &e) {
        std::printf("no device: code %d\n", e.code);
        return e.code == B200_ERR_CUDA ? 0 : 3;
    }
}

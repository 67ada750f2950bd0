// chain_builder.hpp — the op-list fuser behind sl_fused_chain (host side; pure CPU code, no device needed).
//
// The reference's hook for this is custos' `Lazy` graph + `device.optimize()` (examples/chained_perf.rs:114, examples/sine_net.rs:
// 178-233) and the expression closures custos turns into OpenCL source (`to_cl_source`, src/ops2/binary_ew/mod.rs:55-63).  Here a
// chain of element-wise ops recorded from the tape is compiled into two micro-op programs:
//
//   forward   one pass computing the requested values from the chain's leaf buffers (intermediates stay in registers);
//   backward  the tape of the same ops in reverse registration order, as ONE pass: the forward values it needs are recomputed in
//             registers, every `+=` of the reference's grad closures (src/ops.rs:125-126,144-145,163-164,38-41,67-73;
//             src/ops2/binary_ew/grad/cpu_stack.rs:54-59) is performed in the reference's order, leaf gradients are read,
//             accumulated and written back in place.
//
// Both are straight-line SSA first; a linear-scan allocator then maps values to the <= SL_CHAIN_MAX_REGS registers of the
// interpreter (inputs are pinned to r[0..n_in), freed after their last use).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sliced_b200.h"

namespace slh {

class ChainBuilder {
  public:
    // value ids: inputs (leaves) and node results share one numbering, in creation order
    int input();
    int binary(int binop, int lhs, int rhs);                     // SL_ADD / SL_SUB / SL_MUL / SL_DIV
    int unary(int unop, int x, double p0 = 0, double p1 = 0);     // sl_unop forward function
    void no_grad(int v) { nograd_[v] = 1; }                       // a node whose op registers no grad closure (apply_fn, clip, div)
    int n_values() const { return (int)vals_.size(); }
    int n_inputs() const { return n_inputs_; }
    int n_nodes() const { return (int)vals_.size() - n_inputs_; }
    bool is_input(int v) const { return vals_[v].kind == 0; }

    // Forward program: inputs = the leaves in declaration order; output j = value outs[j] (SET).
    // Returns false (and says why) when the chain does not fit the interpreter's limits.
    bool build_forward(const std::vector<int>& outs, sl_chain_prog* prog, std::string* why) const;

    // Backward program.  seeds: values whose gradient arrives from outside (materialised outputs); wrt: leaves that take a gradient.
    // Program inputs, in this order: leaves | one gradient buffer per seed | one gradient buffer per wrt leaf (also outputs, in
    // place: output j = wrt[j], SET of old + contributions in tape order) | then, when write_seed_totals, nothing more: a seed
    // that is itself consumed inside the chain accumulates those contributions too, and every seed's total is what flows on.
    // n_inputs() + seeds.size() + wrt.size() must be <= SL_CHAIN_MAX_INPUTS and wrt.size() <= SL_CHAIN_MAX_OUTPUTS.
    // seed_totals (optional): indices into `seeds` of the seeds that got such an extra output, in output order.
    bool build_backward(const std::vector<int>& seeds, const std::vector<int>& wrt, sl_chain_prog* prog, std::string* why,
                        std::vector<int>* seed_totals = nullptr) const;

  private:
    struct Val {
        int kind;        // 0 input, 1 binary, 2 unary
        int op;          // binop / unop
        int a, b;        // operand value ids
        double p0, p1;
    };
    std::vector<Val> vals_;
    std::vector<char> nograd_;
    int n_inputs_ = 0;

    // straight-line SSA produced by the builders before register allocation
    struct Ssa {
        int op;          // sl_chain_op
        int a, b;        // SSA ids (< 0: unused)
        double p0, p1;
    };
    struct SsaProg {
        int n_in = 0;
        std::vector<Ssa> code;          // value id of code[i] = n_in + i
        std::vector<int> outs;          // SSA ids stored to the outputs
    };
    static bool allocate(const SsaProg& s, sl_chain_prog* prog, std::string* why);
    int emit_value(SsaProg& s, std::vector<int>& ssa_of, int v) const;   // (re)computes value v, memoised in ssa_of
};

}  // namespace slh

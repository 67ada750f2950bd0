// chain_builder.cpp — see chain_builder.hpp.
#include "chain_builder.hpp"

#include <algorithm>
#include <climits>
#include <cstring>

namespace slh {

int ChainBuilder::input() {
    vals_.push_back(Val{0, n_inputs_, -1, -1, 0, 0});   // op = index among the inputs
    nograd_.push_back(0);
    ++n_inputs_;
    return (int)vals_.size() - 1;
}

int ChainBuilder::binary(int binop, int lhs, int rhs) {
    vals_.push_back(Val{1, binop, lhs, rhs, 0, 0});
    nograd_.push_back(0);
    return (int)vals_.size() - 1;
}

int ChainBuilder::unary(int unop, int x, double p0, double p1) {
    vals_.push_back(Val{2, unop, x, -1, p0, p1});
    nograd_.push_back(0);
    return (int)vals_.size() - 1;
}

int ChainBuilder::emit_value(SsaProg& s, std::vector<int>& ssa_of, int v) const {
    if (ssa_of[v] >= 0) return ssa_of[v];
    const Val& x = vals_[v];
    int id;
    if (x.kind == 0) {
        id = x.op;
    } else if (x.kind == 1) {
        const int a = emit_value(s, ssa_of, x.a), b = emit_value(s, ssa_of, x.b);
        s.code.push_back(Ssa{x.op, a, b, 0, 0});
        id = s.n_in + (int)s.code.size() - 1;
    } else {
        const int a = emit_value(s, ssa_of, x.a);
        s.code.push_back(Ssa{SL_CH_UNARY_F + x.op, a, -1, x.p0, x.p1});
        id = s.n_in + (int)s.code.size() - 1;
    }
    return ssa_of[v] = id;
}

bool ChainBuilder::build_forward(const std::vector<int>& outs, sl_chain_prog* prog, std::string* why) const {
    if (outs.empty() || outs.size() > SL_CHAIN_MAX_OUTPUTS) { if (why) *why = "1..4 outputs"; return false; }
    if (n_inputs_ > SL_CHAIN_MAX_INPUTS) { if (why) *why = "too many inputs"; return false; }
    SsaProg s;
    s.n_in = n_inputs_;
    std::vector<int> ssa_of(vals_.size(), -1);
    for (int v : outs) {
        if (v < 0 || v >= (int)vals_.size()) { if (why) *why = "bad output value"; return false; }
        int id = emit_value(s, ssa_of, v);
        if (id < s.n_in) {   // an output that is a bare input: copy it through a register of its own
            s.code.push_back(Ssa{SL_CH_COPY, id, -1, 0, 0});
            id = s.n_in + (int)s.code.size() - 1;
        }
        s.outs.push_back(id);
    }
    return allocate(s, prog, why);
}

bool ChainBuilder::build_backward(const std::vector<int>& seeds, const std::vector<int>& wrt, sl_chain_prog* prog, std::string* why,
                                  std::vector<int>* seed_totals) const {
    const int S = (int)seeds.size(), W = (int)wrt.size();
    if (S < 1) { if (why) *why = "no seed"; return false; }
    if (W < 1 || W > SL_CHAIN_MAX_OUTPUTS) { if (why) *why = "1..4 gradients"; return false; }
    if (n_inputs_ + S + W > SL_CHAIN_MAX_INPUTS) { if (why) *why = "too many inputs (leaves + seeds + gradients)"; return false; }
    SsaProg s;
    s.n_in = n_inputs_ + S + W;
    std::vector<int> ssa_of(vals_.size(), -1), grad(vals_.size(), -1);
    std::vector<char> wanted(vals_.size(), 0), is_seed(vals_.size(), 0), got_internal(vals_.size(), 0);
    for (int i = 0; i < S; ++i) {
        if (seeds[i] < 0 || seeds[i] >= (int)vals_.size() || vals_[seeds[i]].kind == 0) { if (why) *why = "bad seed"; return false; }
        grad[seeds[i]] = n_inputs_ + i;
        is_seed[seeds[i]] = 1;
    }
    for (int i = 0; i < W; ++i) {
        if (wrt[i] < 0 || wrt[i] >= (int)vals_.size() || vals_[wrt[i]].kind != 0) { if (why) *why = "gradients are taken w.r.t. leaves"; return false; }
        grad[wrt[i]] = n_inputs_ + S + i;
        wanted[wrt[i]] = 1;
    }
    for (size_t v = 0; v < vals_.size(); ++v)
        if (vals_[v].kind != 0) wanted[v] = 1;   // intermediates always carry their gradient on (in registers)
    auto push = [&](int op, int a, int b, double p0 = 0, double p1 = 0) {
        s.code.push_back(Ssa{op, a, b, p0, p1});
        return s.n_in + (int)s.code.size() - 1;
    };
    auto accumulate = [&](int v, int c) {   // the reference's `grad[v] += c`, in tape order
        grad[v] = grad[v] < 0 ? c : push(SL_CH_ADD, grad[v], c);
        if (is_seed[v]) got_internal[v] = 1;
    };
    // the ssa ids of forward values refer to LEAF inputs 0..n_inputs_-1, exactly as in the forward program
    for (int v = (int)vals_.size() - 1; v >= 0; --v) {
        const Val& x = vals_[v];
        if (x.kind == 0 || grad[v] < 0 || nograd_[v]) continue;
        const int g = grad[v];
        if (x.kind == 1) {
            const bool wl = wanted[x.a], wr = wanted[x.b];
            int cl = -1, cr = -1;
            switch (x.op) {
            case SL_ADD: cl = g; cr = g; break;                                            // (1, 1)   src/ops.rs:125-126
            case SL_SUB: cl = g; if (wr) cr = push(SL_CH_UNARY_F + SL_UN_NEG, g, -1); break;   // (1, -1)  src/ops.rs:144-145
            case SL_MUL:                                                                   // (r, l)   src/ops.rs:163-164
                if (wl) cl = push(SL_CH_MUL, emit_value(s, ssa_of, x.b), g);
                if (wr) cr = push(SL_CH_MUL, emit_value(s, ssa_of, x.a), g);
                break;
            case SL_DIV: {                                                                 // (1/r, l / -(r*r))   include/sliced_b200.h
                const int rv = emit_value(s, ssa_of, x.b);
                if (wl) cl = push(SL_CH_MUL, push(SL_CH_RDIV_IMM, rv, -1, 1.0), g);
                if (wr) {
                    const int lv = emit_value(s, ssa_of, x.a);
                    const int nrr = push(SL_CH_UNARY_F + SL_UN_NEG, push(SL_CH_MUL, rv, rv), -1);
                    cr = push(SL_CH_MUL, push(SL_CH_DIV, lv, nrr), g);
                }
                break;
            }
            default: if (why) *why = "bad binop"; return false;
            }
            if (wl) accumulate(x.a, cl);   // lhs first, then rhs: binary_ew/grad/cpu_stack.rs:54-59
            if (wr) accumulate(x.b, cr);
        } else if (wanted[x.a]) {          // custos add_unary_grad: xg += f'(x) * og   (src/ops.rs:38-41,67-73)
            const int d = push(SL_CH_UNARY_D + x.op, emit_value(s, ssa_of, x.a), -1, x.p0, x.p1);
            accumulate(x.a, push(SL_CH_MUL, d, g));
        }
    }
    for (int i = 0; i < W; ++i) {
        int id = grad[wrt[i]];
        if (id < s.n_in) id = push(SL_CH_COPY, id, -1);   // untouched: write the old value back
        s.outs.push_back(id);
    }
    // a seed that also fed ops inside the chain: its buffer receives those contributions too (what `.grad()` of that buffer shows
    // in the reference); reported to the caller through prog->n_out > wrt.size(): extra output k = k-th seed with got_internal
    for (int i = 0; i < S; ++i)
        if (got_internal[seeds[i]]) {
            if ((int)s.outs.size() >= SL_CHAIN_MAX_OUTPUTS) { if (why) *why = "too many gradient outputs"; return false; }
            s.outs.push_back(grad[seeds[i]]);
            if (seed_totals) seed_totals->push_back(i);
        }
    return allocate(s, prog, why);
}

// dead-code elimination + linear-scan register allocation; inputs pinned to r[0..n_in), every register reusable after its last read
bool ChainBuilder::allocate(const SsaProg& s, sl_chain_prog* prog, std::string* why) {
    const int n_in = s.n_in, n = (int)s.code.size();
    std::vector<char> used(n_in + n, 0);
    for (int o : s.outs) used[o] = 1;
    for (int i = n - 1; i >= 0; --i) {
        if (!used[n_in + i]) continue;
        if (s.code[i].a >= 0) used[s.code[i].a] = 1;
        if (s.code[i].b >= 0) used[s.code[i].b] = 1;
    }
    std::vector<int> last(n_in + n, -1);
    for (int i = 0; i < n; ++i) {
        if (!used[n_in + i]) continue;
        if (s.code[i].a >= 0) last[s.code[i].a] = i;
        if (s.code[i].b >= 0) last[s.code[i].b] = i;
    }
    for (int o : s.outs) last[o] = INT_MAX;
    std::memset(prog, 0, sizeof(*prog));
    std::vector<int> reg(n_in + n, -1);
    std::vector<int> free_regs;
    int n_regs = n_in;
    for (int r = 0; r < n_in; ++r) {
        reg[r] = r;
        if (last[r] < 0) free_regs.push_back(r);   // an input nobody reads: its register is free from the start
    }
    int k = 0;
    for (int i = 0; i < n; ++i) {
        if (!used[n_in + i]) continue;
        if (k >= SL_CHAIN_MAX_INSTRS) { if (why) *why = "chain too long (instructions)"; return false; }
        const Ssa& c = s.code[i];
        sl_chain_instr& in = prog->instr[k++];
        in.op = (uint8_t)c.op;
        in.a = (uint8_t)(c.a >= 0 ? reg[c.a] : 0);
        in.b = (uint8_t)(c.b >= 0 ? reg[c.b] : 0);
        in.imm0 = c.p0;
        in.imm1 = c.p1;
        // operands read for the last time here give their register back before the destination is chosen
        if (c.a >= 0 && last[c.a] == i) free_regs.push_back(reg[c.a]);
        if (c.b >= 0 && c.b != c.a && last[c.b] == i) free_regs.push_back(reg[c.b]);
        int r;
        if (!free_regs.empty()) {
            auto it = std::min_element(free_regs.begin(), free_regs.end());
            r = *it;
            free_regs.erase(it);
        } else {
            r = n_regs++;
            if (n_regs > SL_CHAIN_MAX_REGS) { if (why) *why = "chain too long (registers)"; return false; }
        }
        reg[n_in + i] = r;
        in.dst = (uint8_t)r;
    }
    prog->n_instr = k;
    prog->n_in = n_in;
    prog->n_regs = n_regs;
    prog->n_out = (int)s.outs.size();
    for (size_t j = 0; j < s.outs.size(); ++j) {
        prog->out_reg[j] = (uint8_t)reg[s.outs[j]];
        prog->out_acc[j] = 0;
    }
    return true;
}

}  // namespace slh

// ---------------------------------------------------------------- C bridge (tests drive the builder from Python, no device needed)
extern "C" {

typedef struct slh_chain slh_chain;

slh_chain* slh_chain_new(void) { return (slh_chain*)new slh::ChainBuilder(); }
void slh_chain_free(slh_chain* c) { delete (slh::ChainBuilder*)c; }
int slh_chain_input(slh_chain* c) { return ((slh::ChainBuilder*)c)->input(); }
int slh_chain_binary(slh_chain* c, int binop, int lhs, int rhs) { return ((slh::ChainBuilder*)c)->binary(binop, lhs, rhs); }
int slh_chain_unary(slh_chain* c, int unop, int x, double p0, double p1) { return ((slh::ChainBuilder*)c)->unary(unop, x, p0, p1); }
void slh_chain_no_grad(slh_chain* c, int v) { ((slh::ChainBuilder*)c)->no_grad(v); }
int slh_chain_build_forward(slh_chain* c, const int* outs, int n_outs, sl_chain_prog* prog) {
    return ((slh::ChainBuilder*)c)->build_forward(std::vector<int>(outs, outs + n_outs), prog, nullptr) ? 0 : -1;
}
int slh_chain_build_backward(slh_chain* c, const int* seeds, int n_seeds, const int* wrt, int n_wrt, sl_chain_prog* prog) {
    return ((slh::ChainBuilder*)c)->build_backward(std::vector<int>(seeds, seeds + n_seeds), std::vector<int>(wrt, wrt + n_wrt), prog, nullptr) ? 0 : -1;
}

}  // extern "C"

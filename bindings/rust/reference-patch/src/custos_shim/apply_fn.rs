//! custos side of the boundary: `ApplyFunction::apply_fn` and `UnaryGrad::add_unary_grad` for `CUDA<Mods>` (custos †; call sites in
//! sliced: src/ops.rs:36,40,65,70-72,423,438; src/matrix.rs:181,186,218,223-225,246-250,255-257).
//!
//! custos hands both an expression CLOSURE `Fn(Resolve<T>) -> impl Eval<T> + MayToCLSource`.  Its OpenCL device evaluates the
//! closure on a marker operand and compiles the resulting source text (`to_cl_source`).  A precompiled CUDA library cannot compile
//! text at run time and does not need to: the same trick — evaluate the closure on a SYMBOLIC operand — yields a micro-op program
//! for `sl_fused_chain` instead of a string.  `SlExpr` below is that symbolic operand: every `Combiner` method custos' expression
//! DSL offers (add / sub / mul / div / pow / neg / exp / ln / tanh / geq / min / max / ...) appends one instruction to the program
//! under construction and returns the register holding the result.  Any expression the DSL can write therefore runs as ONE fused,
//! bandwidth-bound kernel, bit-identical to the op-by-op evaluation (include/sliced_b200.h, `sl_fused_chain`).
use core::cell::RefCell;
use std::rc::Rc;

use custos::{Buffer, OnDropBuffer, Retrieve, Retriever, Shape, CUDA};
use sliced_b200_sys::*;

use crate::cuda_device::{cptr, mptr, SlDevice};

/// program under construction: r[0] = x (and r[1] = out_grad for the gradient form)
pub struct ProgBuilder {
    pub prog: sl_chain_prog,
}

impl ProgBuilder {
    pub fn new(n_in: i32) -> Self {
        let mut prog: sl_chain_prog = unsafe { core::mem::zeroed() };
        prog.n_in = n_in;
        prog.n_regs = n_in;
        Self { prog }
    }
    fn push(&mut self, op: i32, a: u8, b: u8, imm0: f64, imm1: f64) -> u8 {
        let k = self.prog.n_instr as usize;
        assert!(k < SL_CHAIN_MAX_INSTRS && (self.prog.n_regs as usize) < SL_CHAIN_MAX_REGS, "expression too long for one fused chain");
        let dst = self.prog.n_regs as u8;
        self.prog.instr[k] = sl_chain_instr { op: op as u8, dst, a, b, flags: 0, pad_: [0; 3], imm0, imm1 };
        self.prog.n_instr += 1;
        self.prog.n_regs += 1;
        dst
    }
}

/// the symbolic operand handed to the closure
#[derive(Clone)]
pub struct SlExpr {
    b: Rc<RefCell<ProgBuilder>>,
    reg: u8,
}

/// a scalar literal inside an expression (`x.mul(2.)`) or another symbolic value
pub enum Operand {
    Reg(u8),
    Lit(f64),
}
impl From<&SlExpr> for Operand {
    fn from(e: &SlExpr) -> Self {
        Operand::Reg(e.reg)
    }
}
impl From<f64> for Operand {
    fn from(v: f64) -> Self {
        Operand::Lit(v)
    }
}

impl SlExpr {
    fn un(&self, unop: i32, p0: f64, p1: f64) -> SlExpr {
        let reg = self.b.borrow_mut().push(SL_CH_UNARY_F + unop, self.reg, 0, p0, p1);
        SlExpr { b: self.b.clone(), reg }
    }
    fn bin(&self, op: i32, rhs: Operand) -> SlExpr {
        let r = match rhs {
            Operand::Reg(r) => r,
            Operand::Lit(v) => self.b.borrow_mut().push(SL_CH_CONST, 0, 0, v, 0.0),
        };
        let reg = self.b.borrow_mut().push(op, self.reg, r, 0.0, 0.0);
        SlExpr { b: self.b.clone(), reg }
    }
    // ---- the `Combiner` surface custos' closures are written against (one instruction each)
    pub fn add(&self, rhs: impl Into<Operand>) -> SlExpr {
        match rhs.into() {
            Operand::Lit(v) => self.un(SL_UN_ADD_SCALAR, v, 0.0),
            r => self.bin(SL_CH_ADD, r),
        }
    }
    pub fn sub(&self, rhs: impl Into<Operand>) -> SlExpr {
        self.bin(SL_CH_SUB, rhs.into())
    }
    pub fn mul(&self, rhs: impl Into<Operand>) -> SlExpr {
        match rhs.into() {
            Operand::Lit(v) => self.un(SL_UN_MUL_SCALAR, v, 0.0),
            r => self.bin(SL_CH_MUL, r),
        }
    }
    pub fn div(&self, rhs: impl Into<Operand>) -> SlExpr {
        self.bin(SL_CH_DIV, rhs.into())
    }
    pub fn pow(&self, p: f64) -> SlExpr {
        self.un(SL_UN_POW, p, 0.0)
    }
    pub fn neg(&self) -> SlExpr {
        self.un(SL_UN_NEG, 0.0, 0.0)
    }
    pub fn exp(&self) -> SlExpr {
        self.un(SL_UN_EXP, 0.0, 0.0)
    }
    pub fn ln(&self) -> SlExpr {
        self.un(SL_UN_LN, 0.0, 0.0)
    }
    pub fn tanh(&self) -> SlExpr {
        self.un(SL_UN_TANH, 0.0, 0.0)
    }
    /// `x.geq(0.)` — custos evaluates comparisons to 1 / 0 of T (src/matrix.rs:181,186): the relu derivative code is exactly that
    pub fn geq_zero(&self) -> SlExpr {
        let reg = self.b.borrow_mut().push(SL_CH_UNARY_D + SL_UN_RELU, self.reg, 0, 0.0, 0.0);
        SlExpr { b: self.b.clone(), reg }
    }
    pub fn clip(&self, lo: f64, hi: f64) -> SlExpr {
        self.un(SL_UN_CLIP, lo, hi)
    }
}

/// `ApplyFunction::apply_fn`: out[i] = f(x[i])  (SET)
pub fn apply_fn<T: SlDtype, S: Shape, Mods: Retrieve<CUDA<Mods>, T, S>>(
    dev: &CUDA<Mods>, x: &Buffer<T, CUDA<Mods>, S>, f: impl Fn(SlExpr) -> SlExpr,
) -> Buffer<T, CUDA<Mods>, S> {
    let b = Rc::new(RefCell::new(ProgBuilder::new(1)));
    let res = f(SlExpr { b: b.clone(), reg: 0 });
    let mut prog = b.borrow().prog;
    prog.n_out = 1;
    prog.out_reg[0] = res.reg;
    let mut out = dev.retrieve(x.len(), x).unwrap();
    let ins = [cptr(x)];
    let outs = [mptr(&mut out)];
    dev.check(unsafe { sl_fused_chain(dev.ctx(), T::CODE, &prog, ins.as_ptr(), outs.as_ptr(), x.len()) }).unwrap();
    out
}

/// `UnaryGrad::add_unary_grad`: x_grad[i] += g(x[i]) * out_grad[i]  (ACC)
pub fn add_unary_grad<T: SlDtype, S: Shape, Mods: OnDropBuffer>(
    dev: &CUDA<Mods>, x: &Buffer<T, CUDA<Mods>, S>, x_grad: &mut Buffer<T, CUDA<Mods>, S>, out_grad: &Buffer<T, CUDA<Mods>, S>,
    g: impl Fn(SlExpr) -> SlExpr,
) {
    let b = Rc::new(RefCell::new(ProgBuilder::new(2)));
    let d = g(SlExpr { b: b.clone(), reg: 0 });
    let og = SlExpr { b: b.clone(), reg: 1 };
    let contrib = d.mul(&og);
    let mut prog = b.borrow().prog;
    prog.n_out = 1;
    prog.out_reg[0] = contrib.reg;
    prog.out_acc[0] = 1;
    let ins = [cptr(x), cptr(out_grad)];
    let outs = [mptr(x_grad)];
    dev.check(unsafe { sl_fused_chain(dev.ctx(), T::CODE, &prog, ins.as_ptr(), outs.as_ptr(), x.len()) }).unwrap();
}

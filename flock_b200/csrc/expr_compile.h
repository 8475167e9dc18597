// expr_compile.h -- host-side lowering of postfix PhysicalExpr programs to device programs.
#pragma once

#include <string>
#include <vector>

#include "expr_program.h"
#include "internal.h"

namespace fg {

// Input column description the compiler needs (decoupled from Table so it can be unit-tested).
struct ColInfo {
  int dtype;
  std::string name;
  std::string format;
  bool has_nulls = false;  // the column carries validity bytes: comparisons over it are three-valued
};

// Shapes with a hand-specialised kernel (everything else runs the generic term interpreter).
enum FastPredKind {
  FAST_PRED_NONE = 0,
  FAST_PRED_I32_MOD_CMP = 1,  // CAST(i32col AS Int64) % m  CMP  c   (NEXMark q2, planner.rs:122)
  FAST_PRED_I32_CMP = 2       // CAST(i32col AS Int64)      CMP  c   (NEXMark q3 auction side, planner.rs:155)
};
struct FastPred {
  int kind = FAST_PRED_NONE;
  int col = -1;
  int cmp = 0;
  int64_t modulus = 0;
  int64_t rhs = 0;
};

struct CompiledPredicate {
  Predicate prog;
  FastPred fast;
};

enum FastValueKind {
  FAST_VAL_NONE = 0,
  FAST_VAL_I32_TO_F64_MUL = 1  // lit * CAST(i32col AS Float64)   (NEXMark q1, planner.rs:90)
};
struct CompiledValue {
  bool passthrough = false;
  int src_col = -1;
  Chain chain{};
  int dtype = FLOCKGPU_INT64;
  std::string format;  // Arrow format of the result
  bool has_div_by_col = false;
  int fast = FAST_VAL_NONE;
  double fast_lit = 0;
};

Expr tokens_to_expr(const flockgpu_expr* e);
CompiledPredicate compile_predicate(const Expr& e, const std::vector<ColInfo>& cols);
CompiledValue compile_value(const Expr& e, const std::vector<ColInfo>& cols);
// Result type of a value expression without building the chain (used for output schemas).
int infer_dtype(const Expr& e, const std::vector<ColInfo>& cols);
std::string expr_to_string(const Expr& e, const std::vector<ColInfo>& cols);

}  // namespace fg

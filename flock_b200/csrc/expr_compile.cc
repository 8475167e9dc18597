// expr_compile.cc -- lowers postfix PhysicalExpr programs (include/flockgpu.h: flockgpu_expr) to the
// term/chain programs of expr_program.h.
//
// The shapes follow what DataFusion's planner emits for the reference's queries: type coercion is
// already explicit (cast_expr / try_cast_expr nodes, flock/src/tests/data/plan/aggregate.json
// "predicate"; "CAST(auction@0 AS Int64) % 123 = 0", flock/src/distributed_plan/planner.rs:122), so
// both sides of a binary_expr normally carry the same type.  Where a caller of the C ABI omits the
// casts, the usual numeric promotion (Float64 > 64-bit ints > Int32) is applied.
#include "expr_compile.h"

#include <cmath>
#include <cstdio>
#include <functional>

namespace fg {

namespace {

struct Node {
  int op = 0;
  int dtype = 0;  // result type
  int col = -1;
  int64_t i = 0;
  double d = 0;
  std::string s;
  int l = -1, r = -1;
};

bool is_arith(int op) { return op >= FLOCKGPU_OP_ADD && op <= FLOCKGPU_OP_MOD; }
bool is_cmp(int op) { return op >= FLOCKGPU_OP_EQ && op <= FLOCKGPU_OP_GE; }
bool is_literal(const Node& n) { return n.op == FLOCKGPU_OP_LIT_I64 || n.op == FLOCKGPU_OP_LIT_F64 || n.op == FLOCKGPU_OP_LIT_UTF8; }
bool is_intlike(int dt) {
  return dt == FLOCKGPU_INT32 || dt == FLOCKGPU_INT64 || dt == FLOCKGPU_TIMESTAMP || dt == FLOCKGPU_UINT32 || dt == FLOCKGPU_UINT64;
}
bool is_unsigned(int dt) { return dt == FLOCKGPU_UINT32 || dt == FLOCKGPU_UINT64; }
bool is_numeric(int dt) { return is_intlike(dt) || dt == FLOCKGPU_FLOAT64; }

int promote(int a, int b) {
  if (a == b) return a;
  if (a == FLOCKGPU_FLOAT64 || b == FLOCKGPU_FLOAT64) return FLOCKGPU_FLOAT64;
  auto rank = [](int t) {
    switch (t) {
      case FLOCKGPU_INT32: return 1;
      case FLOCKGPU_UINT32: return 2;
      case FLOCKGPU_INT64: case FLOCKGPU_TIMESTAMP: return 3;
      default: return 4;  // UInt64
    }
  };
  return rank(a) >= rank(b) ? a : b;
}

struct Tree {
  std::vector<Node> nodes;
  int root = -1;
};

Tree build_tree(const Expr& e, const std::vector<ColInfo>& cols) {
  Tree t;
  std::vector<int> stack;
  auto pop = [&]() {
    FG_CHECK(!stack.empty(), FLOCKGPU_ERR_INVALID, "expression: stack underflow (malformed postfix program)");
    int v = stack.back();
    stack.pop_back();
    return v;
  };
  for (const ExprTok& tok : e) {
    Node n;
    n.op = tok.op;
    switch (tok.op) {
      case FLOCKGPU_OP_COLUMN:
        FG_CHECK(tok.col >= 0 && tok.col < int(cols.size()), FLOCKGPU_ERR_INVALID, "expression: column index %d out of range (%zu columns)",
                 tok.col, cols.size());
        n.col = tok.col;
        n.dtype = cols[tok.col].dtype;
        break;
      case FLOCKGPU_OP_LIT_I64:
        n.i = tok.i64;
        n.dtype = (tok.dtype == FLOCKGPU_INT32 || tok.dtype == FLOCKGPU_UINT32 || tok.dtype == FLOCKGPU_UINT64 ||
                   tok.dtype == FLOCKGPU_TIMESTAMP)
                      ? tok.dtype
                      : FLOCKGPU_INT64;
        break;
      case FLOCKGPU_OP_LIT_F64:
        n.d = tok.f64;
        n.dtype = FLOCKGPU_FLOAT64;
        break;
      case FLOCKGPU_OP_LIT_UTF8:
        n.s = tok.str;
        n.dtype = FLOCKGPU_UTF8;
        break;
      case FLOCKGPU_OP_CAST:
        n.l = pop();
        n.dtype = tok.dtype;
        FG_CHECK(is_numeric(n.dtype) && is_numeric(t.nodes[n.l].dtype), FLOCKGPU_ERR_UNSUPPORTED, "expression: CAST from %s to %s",
                 dtype_name(t.nodes[n.l].dtype), dtype_name(n.dtype));
        break;
      case FLOCKGPU_OP_NOT:
        n.l = pop();
        n.dtype = FLOCKGPU_BOOL;
        FG_CHECK(t.nodes[n.l].dtype == FLOCKGPU_BOOL, FLOCKGPU_ERR_INVALID, "expression: NOT of a non-boolean");
        break;
      default:
        if (is_arith(tok.op) || is_cmp(tok.op) || tok.op == FLOCKGPU_OP_AND || tok.op == FLOCKGPU_OP_OR) {
          n.r = pop();
          n.l = pop();
          int lt = t.nodes[n.l].dtype, rt = t.nodes[n.r].dtype;
          if (is_arith(tok.op)) {
            FG_CHECK(is_numeric(lt) && is_numeric(rt), FLOCKGPU_ERR_UNSUPPORTED, "expression: arithmetic on %s and %s", dtype_name(lt),
                     dtype_name(rt));
            n.dtype = promote(lt, rt);
          } else if (is_cmp(tok.op)) {
            FG_CHECK((is_numeric(lt) && is_numeric(rt)) || (lt == FLOCKGPU_UTF8 && rt == FLOCKGPU_UTF8), FLOCKGPU_ERR_UNSUPPORTED,
                     "expression: comparison of %s with %s", dtype_name(lt), dtype_name(rt));
            n.dtype = FLOCKGPU_BOOL;
          } else {
            FG_CHECK(lt == FLOCKGPU_BOOL && rt == FLOCKGPU_BOOL, FLOCKGPU_ERR_INVALID, "expression: AND/OR of non-boolean operands");
            n.dtype = FLOCKGPU_BOOL;
          }
        } else {
          fail(FLOCKGPU_ERR_INVALID, "expression: unknown opcode %d", tok.op);
        }
    }
    t.nodes.push_back(n);
    stack.push_back(int(t.nodes.size()) - 1);
  }
  FG_CHECK(stack.size() == 1, FLOCKGPU_ERR_INVALID, "expression: %zu values left on the stack (malformed postfix program)", stack.size());
  t.root = stack[0];
  return t;
}

// Literal value converted to the canonical representation of `dtype`.
Val literal_as(const Node& n, int dtype) {
  Val v;
  v.i = 0;
  if (dtype == FLOCKGPU_FLOAT64) {
    v.d = n.op == FLOCKGPU_OP_LIT_F64 ? n.d : (is_unsigned(n.dtype) ? double(uint64_t(n.i)) : double(n.i));
  } else {
    int64_t x = n.op == FLOCKGPU_OP_LIT_F64 ? int64_t(n.d) : n.i;
    if (dtype == FLOCKGPU_INT32) x = int64_t(int32_t(uint32_t(uint64_t(x))));
    if (dtype == FLOCKGPU_UINT32) x = int64_t(uint64_t(uint32_t(uint64_t(x))));
    v.i = x;
  }
  return v;
}

struct ChainBuilder {
  const Tree& t;
  bool div_by_col = false;
  explicit ChainBuilder(const Tree& tree) : t(tree) {}

  static void push(Chain& c, ChainStep st) {
    FG_CHECK(c.n_steps < MAX_STEPS, FLOCKGPU_ERR_UNSUPPORTED, "expression: more than %d arithmetic steps in one chain", MAX_STEPS);
    c.steps[c.n_steps++] = st;
  }
  static void push_unary(Chain& c, int op) {
    ChainStep st{};
    st.op = int16_t(op);
    st.src = 2;
    push(c, st);
  }
  // steps converting the accumulator from `from` to `to`
  static void convert(Chain& c, int from, int to) {
    if (from == to) return;
    if (to == FLOCKGPU_FLOAT64) {
      push_unary(c, is_unsigned(from) ? C_U2F : C_I2F);
    } else if (from == FLOCKGPU_FLOAT64) {
      push_unary(c, C_F2I);
      if (to == FLOCKGPU_INT32) push_unary(c, C_WRAP_I32);
      if (to == FLOCKGPU_UINT32) push_unary(c, C_WRAP_U32);
    } else {
      if (to == FLOCKGPU_INT32) push_unary(c, C_WRAP_I32);
      else if (to == FLOCKGPU_UINT32) push_unary(c, C_WRAP_U32);
      else if (from == FLOCKGPU_UINT32 || from == FLOCKGPU_INT32) { /* canonical form already 64-bit */ }
    }
  }

  // A "simple" operand: literal, column, or CAST of either that needs at most a load-time conversion.
  // Returns false if `idx` is not simple.  On success fills (is_col, col, cvt, lit) for result type T.
  bool simple_operand(int idx, int T, bool* is_col, int* col, int8_t* cvt, Val* lit) const {
    const Node* n = &t.nodes[idx];
    int src_dtype = n->dtype;
    while (n->op == FLOCKGPU_OP_CAST) {
      const Node* inner = &t.nodes[n->l];
      // only widening / to-float casts keep the operand simple
      bool widening = (n->dtype == FLOCKGPU_FLOAT64) || (is_intlike(n->dtype) && is_intlike(inner->dtype) && dtype_width(n->dtype) >= dtype_width(inner->dtype));
      if (!widening && !is_literal(*inner)) return false;
      if (inner->dtype == FLOCKGPU_FLOAT64 && n->dtype != FLOCKGPU_FLOAT64 && !is_literal(*inner)) return false;
      src_dtype = n->dtype;
      n = inner;
    }
    (void)src_dtype;
    if (is_literal(*n)) {
      if (n->op == FLOCKGPU_OP_LIT_UTF8) return false;
      // fold the cast chain on the host: evaluate through each cast level
      std::function<Val(int, int*)> fold = [&](int i, int* dt) -> Val {
        const Node& m = t.nodes[i];
        if (m.op == FLOCKGPU_OP_CAST) {
          int inner_dt;
          Val v = fold(m.l, &inner_dt);
          Node tmp;
          tmp.dtype = inner_dt;
          if (inner_dt == FLOCKGPU_FLOAT64) { tmp.op = FLOCKGPU_OP_LIT_F64; tmp.d = v.d; }
          else { tmp.op = FLOCKGPU_OP_LIT_I64; tmp.i = v.i; }
          *dt = m.dtype;
          return literal_as(tmp, m.dtype);
        }
        *dt = m.dtype;
        return literal_as(m, m.dtype);
      };
      int dt;
      Val v = fold(idx, &dt);
      Node tmp;
      tmp.dtype = dt;
      if (dt == FLOCKGPU_FLOAT64) { tmp.op = FLOCKGPU_OP_LIT_F64; tmp.d = v.d; }
      else { tmp.op = FLOCKGPU_OP_LIT_I64; tmp.i = v.i; }
      *is_col = false;
      *lit = literal_as(tmp, T);
      *cvt = CVT_NONE;
      return true;
    }
    if (n->op == FLOCKGPU_OP_COLUMN) {
      *is_col = true;
      *col = n->col;
      *cvt = CVT_NONE;
      if (T == FLOCKGPU_FLOAT64 && n->dtype != FLOCKGPU_FLOAT64) *cvt = is_unsigned(n->dtype) ? CVT_U2F : CVT_I2F;
      if (T != FLOCKGPU_FLOAT64 && n->dtype == FLOCKGPU_FLOAT64) return false;
      return true;
    }
    return false;
  }

  static int arith_opcode(int op, int T, bool reversed) {
    bool f = T == FLOCKGPU_FLOAT64, u = is_unsigned(T);
    switch (op) {
      case FLOCKGPU_OP_ADD: return f ? C_ADD_F : C_ADD_I;
      case FLOCKGPU_OP_MUL: return f ? C_MUL_F : C_MUL_I;
      case FLOCKGPU_OP_SUB: return f ? (reversed ? C_RSUB_F : C_SUB_F) : (reversed ? C_RSUB_I : C_SUB_I);
      case FLOCKGPU_OP_DIV: return f ? (reversed ? C_RDIV_F : C_DIV_F) : u ? (reversed ? C_RDIV_U : C_DIV_U) : (reversed ? C_RDIV_I : C_DIV_I);
      case FLOCKGPU_OP_MOD:
        FG_CHECK(!f, FLOCKGPU_ERR_UNSUPPORTED, "expression: %% on Float64");
        return u ? (reversed ? C_RMOD_U : C_MOD_U) : (reversed ? C_RMOD_I : C_MOD_I);
    }
    fail(FLOCKGPU_ERR_INVALID, "expression: bad arithmetic opcode %d", op);
  }

  // Builds the chain computing node `idx`; the accumulator ends in the canonical form of node.dtype.
  Chain build(int idx) {
    const Node& n = t.nodes[idx];
    Chain c{};
    c.start_col = -1;
    c.n_steps = 0;
    c.start_lit.i = 0;
    if (n.op == FLOCKGPU_OP_COLUMN) {
      FG_CHECK(n.dtype != FLOCKGPU_UTF8, FLOCKGPU_ERR_UNSUPPORTED, "expression: Utf8 column in an arithmetic context");
      c.start_col = n.col;
      return c;
    }
    if (is_literal(n)) {
      FG_CHECK(n.op != FLOCKGPU_OP_LIT_UTF8, FLOCKGPU_ERR_UNSUPPORTED, "expression: Utf8 literal in an arithmetic context");
      c.start_lit = literal_as(n, n.dtype);
      return c;
    }
    if (n.op == FLOCKGPU_OP_CAST) {
      c = build(n.l);
      convert(c, t.nodes[n.l].dtype, n.dtype);
      return c;
    }
    FG_CHECK(is_arith(n.op), FLOCKGPU_ERR_UNSUPPORTED, "expression: boolean sub-expression used as a value");
    int T = n.dtype;
    bool is_col = false;
    int col = -1;
    int8_t cvt = CVT_NONE;
    Val lit;
    lit.i = 0;
    bool reversed = false;
    if (simple_operand(n.r, T, &is_col, &col, &cvt, &lit)) {
      c = build(n.l);
      convert(c, t.nodes[n.l].dtype, T);
    } else if (simple_operand(n.l, T, &is_col, &col, &cvt, &lit)) {
      c = build(n.r);
      convert(c, t.nodes[n.r].dtype, T);
      reversed = true;
    } else {
      fail(FLOCKGPU_ERR_UNSUPPORTED,
           "expression: both operands of an arithmetic node are compound; the GPU path evaluates left-deep chains only");
    }
    ChainStep st{};
    st.op = int16_t(arith_opcode(n.op, T, reversed));
    st.src = is_col ? 1 : 0;
    st.cvt = cvt;
    st.col = col;
    st.lit = lit;
    bool divides = (n.op == FLOCKGPU_OP_DIV || n.op == FLOCKGPU_OP_MOD) && T != FLOCKGPU_FLOAT64;
    if (divides) {
      if (!reversed && !is_col)
        FG_CHECK(lit.i != 0, FLOCKGPU_ERR_EXECUTION, "Divide by zero");  // DataFusion's message for a zero divisor
      if (is_col || reversed) div_by_col = true;
    }
    push(c, st);
    if (T == FLOCKGPU_INT32) push_unary(c, C_WRAP_I32);
    if (T == FLOCKGPU_UINT32) push_unary(c, C_WRAP_U32);
    return c;
  }
};

int cmp_domain(int lt, int rt) {
  if (lt == FLOCKGPU_FLOAT64 || rt == FLOCKGPU_FLOAT64) return DOM_F64;
  if (is_unsigned(lt) && is_unsigned(rt)) return DOM_U64;
  if (is_unsigned(lt) || is_unsigned(rt)) return DOM_U64;  // DataFusion coerces both sides first
  return DOM_I64;
}

int swap_cmp(int cmp) {
  switch (cmp) {
    case FLOCKGPU_OP_LT: return FLOCKGPU_OP_GT;
    case FLOCKGPU_OP_LE: return FLOCKGPU_OP_GE;
    case FLOCKGPU_OP_GT: return FLOCKGPU_OP_LT;
    case FLOCKGPU_OP_GE: return FLOCKGPU_OP_LE;
    default: return cmp;
  }
}

}  // namespace

Expr tokens_to_expr(const flockgpu_expr* e) {
  Expr out;
  FG_CHECK(e && e->tokens && e->n_tokens > 0, FLOCKGPU_ERR_INVALID, "expression: empty program");
  for (int i = 0; i < e->n_tokens; ++i) {
    const flockgpu_expr_token& t = e->tokens[i];
    ExprTok k;
    k.op = t.op;
    k.dtype = t.dtype;
    k.col = t.col;
    k.i64 = t.i64;
    k.f64 = t.f64;
    if (t.op == FLOCKGPU_OP_LIT_UTF8) {
      FG_CHECK(t.str_len >= 0 && (t.str || t.str_len == 0), FLOCKGPU_ERR_INVALID, "expression: bad Utf8 literal");
      k.str.assign(t.str ? t.str : "", size_t(t.str_len));
    }
    out.push_back(std::move(k));
  }
  return out;
}

int infer_dtype(const Expr& e, const std::vector<ColInfo>& cols) {
  Tree t = build_tree(e, cols);
  return t.nodes[t.root].dtype;
}

CompiledValue compile_value(const Expr& e, const std::vector<ColInfo>& cols) {
  Tree t = build_tree(e, cols);
  const Node& root = t.nodes[t.root];
  CompiledValue out;
  out.dtype = root.dtype;
  FG_CHECK(root.dtype != FLOCKGPU_BOOL, FLOCKGPU_ERR_UNSUPPORTED, "projection: boolean-valued output columns are not supported");
  if (root.op == FLOCKGPU_OP_COLUMN) {
    out.passthrough = true;
    out.src_col = root.col;
    out.format = cols[root.col].format;
    return out;
  }
  FG_CHECK(root.dtype != FLOCKGPU_UTF8, FLOCKGPU_ERR_UNSUPPORTED, "projection: computed Utf8 columns are not supported");
  ChainBuilder b(t);
  out.chain = b.build(t.root);
  out.has_div_by_col = b.div_by_col;
  out.format = default_format(root.dtype);
  if (root.dtype == FLOCKGPU_TIMESTAMP) {
    // keep the unit/timezone of the first timestamp column the expression touches
    for (const Node& n : t.nodes)
      if (n.op == FLOCKGPU_OP_COLUMN && n.dtype == FLOCKGPU_TIMESTAMP) {
        out.format = cols[n.col].format;
        break;
      }
  }
  // fast shape: lit * CAST(i32 AS f64)
  const Chain& c = out.chain;
  if (root.dtype == FLOCKGPU_FLOAT64 && c.start_col >= 0 && cols[c.start_col].dtype == FLOCKGPU_INT32 && c.n_steps == 2 &&
      c.steps[0].op == C_I2F && c.steps[1].op == C_MUL_F && c.steps[1].src == 0) {
    out.fast = FAST_VAL_I32_TO_F64_MUL;
    out.fast_lit = c.steps[1].lit.d;
  }
  return out;
}

CompiledPredicate compile_predicate(const Expr& e, const std::vector<ColInfo>& cols) {
  Tree t = build_tree(e, cols);
  FG_CHECK(t.nodes[t.root].dtype == FLOCKGPU_BOOL, FLOCKGPU_ERR_INVALID, "filter: predicate is not boolean");
  CompiledPredicate out;
  memset(&out.prog, 0, sizeof out.prog);
  Predicate& p = out.prog;
  ChainBuilder b(t);
  int pool_used = 0;
  std::vector<int> term_of_node(t.nodes.size(), -1);

  std::function<void(int)> collect = [&](int idx) {
    const Node& n = t.nodes[idx];
    if (n.op == FLOCKGPU_OP_AND || n.op == FLOCKGPU_OP_OR) {
      collect(n.l);
      collect(n.r);
      return;
    }
    if (n.op == FLOCKGPU_OP_NOT) {
      collect(n.l);
      return;
    }
    FG_CHECK(is_cmp(n.op), FLOCKGPU_ERR_UNSUPPORTED, "filter: boolean leaf is not a comparison (opcode %d)", n.op);
    FG_CHECK(p.n_terms < MAX_TERMS, FLOCKGPU_ERR_UNSUPPORTED, "filter: more than %d comparison terms", MAX_TERMS);
    Term& term = p.terms[p.n_terms];
    term.cmp = n.op;
    const Node& l = t.nodes[n.l];
    const Node& r = t.nodes[n.r];
    if (l.dtype == FLOCKGPU_UTF8) {
      term.domain = DOM_UTF8;
      const Node* colside = &l;
      const Node* other = &r;
      if (l.op != FLOCKGPU_OP_COLUMN) {
        colside = &r;
        other = &l;
        term.cmp = swap_cmp(n.op);
      }
      FG_CHECK(colside->op == FLOCKGPU_OP_COLUMN, FLOCKGPU_ERR_UNSUPPORTED, "filter: Utf8 comparison needs a column operand");
      term.lhs_col = colside->col;
      if (other->op == FLOCKGPU_OP_COLUMN) {
        term.rhs_col = other->col;
      } else {
        FG_CHECK(other->op == FLOCKGPU_OP_LIT_UTF8, FLOCKGPU_ERR_UNSUPPORTED, "filter: Utf8 comparison with a computed operand");
        FG_CHECK(pool_used + int(other->s.size()) <= STRPOOL_BYTES, FLOCKGPU_ERR_UNSUPPORTED, "filter: Utf8 literals exceed %d bytes", STRPOOL_BYTES);
        term.rhs_col = -1;
        term.lit_off = pool_used;
        term.lit_len = int(other->s.size());
        memcpy(p.strpool + pool_used, other->s.data(), other->s.size());
        pool_used += int(other->s.size());
      }
    } else {
      term.domain = cmp_domain(l.dtype, r.dtype);
      term.lhs = b.build(n.l);
      term.rhs = b.build(n.r);
      if (term.domain == DOM_F64) {
        ChainBuilder::convert(term.lhs, l.dtype, FLOCKGPU_FLOAT64);
        ChainBuilder::convert(term.rhs, r.dtype, FLOCKGPU_FLOAT64);
      }
    }
    // which input columns make this comparison NULL
    {
      std::function<void(int)> cols_of = [&](int i) {
        const Node& x = t.nodes[i];
        if (x.op == FLOCKGPU_OP_COLUMN) {
          if (x.col >= 0 && x.col < 32 && cols[x.col].has_nulls) term.null_cols |= 1u << x.col;
          return;
        }
        if (x.l >= 0) cols_of(x.l);
        if (x.r >= 0) cols_of(x.r);
      };
      term.null_cols = 0;
      cols_of(n.l);
      cols_of(n.r);
    }
    term_of_node[idx] = p.n_terms++;
  };
  collect(t.root);
  p.has_div_by_col = b.div_by_col ? 1 : 0;
  bool any_null = false;
  for (int i = 0; i < p.n_terms; ++i) any_null |= p.terms[i].null_cols != 0;

  if (!any_null) {
    // truth table over the term bits
    std::function<bool(int, unsigned)> truth = [&](int idx, unsigned bits) -> bool {
      const Node& n = t.nodes[idx];
      if (n.op == FLOCKGPU_OP_AND) return truth(n.l, bits) && truth(n.r, bits);
      if (n.op == FLOCKGPU_OP_OR) return truth(n.l, bits) || truth(n.r, bits);
      if (n.op == FLOCKGPU_OP_NOT) return !truth(n.l, bits);
      return (bits >> term_of_node[idx]) & 1u;
    };
    for (unsigned bits = 0; bits < (1u << p.n_terms); ++bits)
      if (truth(t.root, bits)) p.lut[bits >> 5] |= 1u << (bits & 31);
  } else {
    // SQL three-valued logic (strong Kleene): 1 = TRUE, 0 = FALSE, 2 = NULL; the row is kept iff the root is TRUE
    FG_CHECK(p.n_terms <= 4, FLOCKGPU_ERR_UNSUPPORTED, "filter: more than 4 comparison terms over columns with NULLs");
    p.kleene = 1;
    std::function<int(int, unsigned)> k3 = [&](int idx, unsigned bits) -> int {
      const Node& n = t.nodes[idx];
      if (n.op == FLOCKGPU_OP_AND) {
        const int a = k3(n.l, bits), c = k3(n.r, bits);
        return (a == 0 || c == 0) ? 0 : (a == 2 || c == 2) ? 2 : 1;
      }
      if (n.op == FLOCKGPU_OP_OR) {
        const int a = k3(n.l, bits), c = k3(n.r, bits);
        return (a == 1 || c == 1) ? 1 : (a == 2 || c == 2) ? 2 : 0;
      }
      if (n.op == FLOCKGPU_OP_NOT) {
        const int a = k3(n.l, bits);
        return a == 2 ? 2 : 1 - a;
      }
      const int tix = term_of_node[idx];
      return ((bits >> (4 + tix)) & 1u) ? 2 : int((bits >> tix) & 1u);
    };
    for (unsigned bits = 0; bits < 256; ++bits)
      if (k3(t.root, bits) == 1) p.lut[bits >> 5] |= 1u << (bits & 31);
    return out;  // no vectorised shape reads validity
  }

  // fast shapes: one term over one Int32 column against literals
  // (a single term has a 2-entry truth table: 0b10 = the term itself, 0b01 = NOT term; constants keep the interpreter)
  const unsigned table1 = p.lut[0] & 3u;
  auto negate_cmp = [](int cmp) {
    switch (cmp) {
      case FLOCKGPU_OP_EQ: return int(FLOCKGPU_OP_NE);
      case FLOCKGPU_OP_NE: return int(FLOCKGPU_OP_EQ);
      case FLOCKGPU_OP_LT: return int(FLOCKGPU_OP_GE);
      case FLOCKGPU_OP_LE: return int(FLOCKGPU_OP_GT);
      case FLOCKGPU_OP_GT: return int(FLOCKGPU_OP_LE);
      default: return int(FLOCKGPU_OP_LT);
    }
  };
  if (p.n_terms == 1 && p.terms[0].domain == DOM_I64 && (table1 == 2u || table1 == 1u)) {
    const Term& term = p.terms[0];
    const Chain& l = term.lhs;
    const Chain& r = term.rhs;
    if (r.start_col < 0 && r.n_steps == 0 && l.start_col >= 0 && cols[l.start_col].dtype == FLOCKGPU_INT32) {
      if (l.n_steps == 0) {
        out.fast.kind = FAST_PRED_I32_CMP;
        out.fast.col = l.start_col;
        out.fast.cmp = term.cmp;
        out.fast.rhs = r.start_lit.i;
      } else if (l.n_steps == 1 && l.steps[0].op == C_MOD_I && l.steps[0].src == 0) {
        int64_t m = l.steps[0].lit.i;
        int64_t am = m < 0 ? -m : m;
        if (am > 0 && am < (int64_t(1) << 31)) {
          out.fast.kind = FAST_PRED_I32_MOD_CMP;
          out.fast.col = l.start_col;
          out.fast.cmp = term.cmp;
          out.fast.modulus = am;  // x % m == x % |m| (sign follows the dividend)
          out.fast.rhs = r.start_lit.i;
        }
      }
    }
    if (out.fast.kind != FAST_PRED_NONE && table1 == 1u) out.fast.cmp = negate_cmp(out.fast.cmp);
  }
  return out;
}

std::string expr_to_string(const Expr& e, const std::vector<ColInfo>& cols) {
  std::vector<std::string> st;
  auto opname = [](int op) -> const char* {
    switch (op) {
      case FLOCKGPU_OP_ADD: return "+"; case FLOCKGPU_OP_SUB: return "-"; case FLOCKGPU_OP_MUL: return "*";
      case FLOCKGPU_OP_DIV: return "/"; case FLOCKGPU_OP_MOD: return "%"; case FLOCKGPU_OP_EQ: return "=";
      case FLOCKGPU_OP_NE: return "!="; case FLOCKGPU_OP_LT: return "<"; case FLOCKGPU_OP_LE: return "<=";
      case FLOCKGPU_OP_GT: return ">"; case FLOCKGPU_OP_GE: return ">="; case FLOCKGPU_OP_AND: return "AND";
      case FLOCKGPU_OP_OR: return "OR"; default: return "?";
    }
  };
  char buf[64];
  for (const ExprTok& t : e) {
    switch (t.op) {
      case FLOCKGPU_OP_COLUMN:
        st.push_back((t.col >= 0 && t.col < int(cols.size()) ? cols[t.col].name : std::string("?")) + "@" + std::to_string(t.col));
        break;
      case FLOCKGPU_OP_LIT_I64: st.push_back(std::to_string(t.i64)); break;
      case FLOCKGPU_OP_LIT_F64: snprintf(buf, sizeof buf, "%.17g", t.f64); st.push_back(buf); break;
      case FLOCKGPU_OP_LIT_UTF8: st.push_back(t.str); break;
      case FLOCKGPU_OP_CAST: {
        if (st.empty()) return "<malformed>";
        std::string a = st.back(); st.pop_back();
        st.push_back("CAST(" + a + " AS " + dtype_name(t.dtype) + ")");
        break;
      }
      case FLOCKGPU_OP_NOT: {
        if (st.empty()) return "<malformed>";
        std::string a = st.back(); st.pop_back();
        st.push_back("NOT " + a);
        break;
      }
      default: {
        if (st.size() < 2) return "<malformed>";
        std::string r = st.back(); st.pop_back();
        std::string l = st.back(); st.pop_back();
        st.push_back(l + " " + opname(t.op) + " " + r);
      }
    }
  }
  return st.empty() ? "" : st.back();
}

}  // namespace fg

// gm_score.h -- the score program's pre-screen (include/gpumotif_score.h): an
// interpreter of the subset of src/score.c's bytecode that decides REJECT, with the
// reference's semantics instruction by instruction (RM_score, src/score.c:608-790;
// the do_* functions :1130-3090).  Plain C++ that compiles for the device (sink_pass,
// gm_kernel.cuh) and for the host (gm_post.cpp: gm_score_prescreen, the CPU check of
// the same code).  Whatever it cannot be sure of KEEPS the candidate.
#pragma once

#include "gpumotif_plan.h"
#include "gpumotif_score.h"

#ifdef __CUDACC__
#define GM_HD __host__ __device__
#else
#define GM_HD
#endif

namespace gm {

enum { SC_KEEP = 0, SC_REJECT = 1 };

// A value on the program's stack or in a variable.  Strings are references: so >= 0
// = sl characters of the searched strand from offset so on; so < 0 = sl characters of
// the program's string pool from offset -so - 1 on.
struct ScVal {
	int t;
	int i;
	double d;
	int so, sl;
};

#define GM_SC_STACK 32
#define GM_SC_STEPS 20000

// Env supplies the candidate:
//   int ch(int pos)        character of the searched strand at absolute offset pos as
//                          fm_sbuf holds it, or -1 when the device cannot tell
//   int off(int d), len(int d), mpr(int d), mm(int d)   element d of the plan
//   int comp(), pos(), mlen(), slen()                   COMP, POS, LEN, SLEN
//   const gm_elem_t &elem(int d); const gm_pairset_t &pairset(int i)
template <class Env>
GM_HD int sc_char(const gm_score_t &sc, Env &env, const ScVal &v, int k)
{
	return v.so >= 0 ? env.ch(v.so + k) : (int)(unsigned char)sc.str[-v.so - 1 + k];
}

// strcmp of two string values: -1 / 0 / 1, or 2 when a character is not known
template <class Env>
GM_HD int sc_strcmp(const gm_score_t &sc, Env &env, const ScVal &a, const ScVal &b)
{
	const int n = a.sl < b.sl ? a.sl : b.sl;
	for (int k = 0; k < n; k++) {
		const int x = sc_char(sc, env, a, k), y = sc_char(sc, env, b, k);
		if (x < 0 || y < 0)
			return 2;
		if (x != y)
			return x < y ? -1 : 1;
	}
	return a.sl == b.sl ? 0 : a.sl < b.sl ? -1 : 1;
}

template <class Env>
GM_HD int score_eval(const gm_score_t &sc, Env &env)
{
	ScVal mem[GM_SC_STACK];
	ScVal var[GM_SC_MAX_VAR];
	int estk[20]; // ESTK_SIZE (the reference writes past it without a check: kept)
	int sp = -1, mp = -1, esp = -1, pc = 0;
	if (!sc.present || sc.n_inst <= 0)
		return SC_KEEP;
	for (int v = 0; v < sc.n_var; v++) {
		var[v].t = sc.var[v].type;
		var[v].i = sc.var[v].ival;
		var[v].d = sc.var[v].dval;
		var[v].so = -sc.var[v].ival - 1;
		var[v].sl = 0;
		if (var[v].t == GM_T_STRING)
			for (const char *p = sc.str + sc.var[v].ival; *p; p++)
				var[v].sl++;
	}
#ifdef GM_SCORE_TRACE
#define SC_GIVE_UP() (GM_SCORE_TRACE(pc - 1, __LINE__), SC_KEEP)
#else
#define SC_GIVE_UP() SC_KEEP
#endif
#define SC_NEED(c)                \
	do {                          \
		if (!(c))                 \
			return SC_GIVE_UP();  \
	} while (0)
	for (int steps = 0; steps < GM_SC_STEPS; steps++) {
		SC_NEED(pc >= 0 && pc < sc.n_inst);
		const gm_sc_inst_t in = sc.inst[pc++];
		switch (in.op) {
		case GM_OP_NOOP:
			break;
		case GM_OP_RJCT:
			return SC_REJECT;
		case GM_OP_MRK:
			SC_NEED(sp + 1 < GM_SC_STACK);
			sp++;
			mem[sp].t = GM_T_INT;
			mem[sp].i = mp;
			mp = sp;
			break;
		case GM_OP_CLS:
			sp = mp = -1;
			break;
		case GM_OP_LDA:
			SC_NEED(sp + 1 < GM_SC_STACK && in.a >= 0 && in.a < sc.n_var);
			sp++;
			mem[sp].t = GM_T_IDENT;
			mem[sp].i = in.a;
			break;
		case GM_OP_LOD: {
			SC_NEED(sp + 1 < GM_SC_STACK && in.a >= 0 && in.a < sc.n_var);
			const int sv = sc.var[in.a].special;
			sp++;
			if (sv == GM_SV_COMP || sv == GM_SV_POS || sv == GM_SV_LEN || sv == GM_SV_SLEN) {
				mem[sp].t = GM_T_INT;
				mem[sp].i = sv == GM_SV_COMP ? env.comp() : sv == GM_SV_POS ? env.pos() : sv == GM_SV_LEN ? env.mlen() : env.slen();
				break;
			}
			SC_NEED(sv == GM_SV_NONE);
			const ScVal &v = var[in.a];
			SC_NEED(v.t == GM_T_INT || v.t == GM_T_FLOAT || v.t == GM_T_STRING);
			mem[sp] = v;
			break;
		}
		case GM_OP_LDC:
			SC_NEED(sp + 1 < GM_SC_STACK);
			sp++;
			mem[sp].t = in.vtype;
			if (in.vtype == GM_T_INT)
				mem[sp].i = in.a;
			else if (in.vtype == GM_T_FLOAT) {
				SC_NEED(in.a >= 0 && in.a < sc.n_dbl);
				mem[sp].d = sc.dbl[in.a];
			} else if (in.vtype == GM_T_STRING) {
				SC_NEED(in.a >= 0 && in.a < sc.n_str);
				mem[sp].so = -in.a - 1;
				mem[sp].sl = 0;
				for (const char *p = sc.str + in.a; *p; p++)
					mem[sp].sl++;
			} else if (in.vtype == GM_T_POS) {
				// `$`: rm_descr[ estk[esp] ].s_matchlen (:2217-2221) -- the index is one of
				// rm_xdescr, so the two only agree without an explicit left context
				SC_NEED(esp >= 0 && estk[esp] >= 0 && estk[esp] < sc.n_xel && sc.xel[estk[esp]].elem == estk[esp]);
				mem[sp].t = GM_T_INT;
				mem[sp].i = env.len(estk[esp]);
			} else
				mem[sp].t = GM_T_UNKNOWN;
			break;
		case GM_OP_STO: {
			SC_NEED(sp >= 1 && mem[sp - 1].t == GM_T_IDENT);
			const ScVal top = mem[sp];
			sp--;
			const int vi = mem[sp].i;
			SC_NEED(vi >= 0 && vi < sc.n_var && sc.var[vi].special == GM_SV_NONE);
			ScVal &v = var[vi];
			if (v.t == GM_T_UNDEF && (top.t == GM_T_INT || top.t == GM_T_FLOAT || top.t == GM_T_STRING))
				v = top;
			else if (v.t == GM_T_INT && top.t == GM_T_INT)
				v.i = top.i;
			else if (v.t == GM_T_INT && top.t == GM_T_FLOAT) {
				SC_NEED(top.d > -2e9 && top.d < 2e9);
				v.i = (int)top.d;
			} else if (v.t == GM_T_FLOAT && top.t == GM_T_INT)
				v.d = top.i;
			else if (v.t == GM_T_FLOAT && top.t == GM_T_FLOAT)
				v.d = top.d;
			else if (v.t == GM_T_STRING && top.t == GM_T_STRING) {
				v.so = top.so;
				v.sl = top.sl;
			} else
				return SC_KEEP; // "type mismatch" (:2298-2302)
			mem[sp].t = GM_T_UNKNOWN; // what the reference leaves there is not a value
			break;
		}
		case GM_OP_AND:
		case GM_OP_IOR: {
			SC_NEED(sp >= 0 && mem[sp].t == GM_T_INT);
			const int rv = mem[sp].i != 0;
			mem[sp].i = rv;
			if (in.op == GM_OP_AND ? !rv : rv)
				pc = in.a;
			break;
		}
		case GM_OP_NOT:
			SC_NEED(sp >= 0 && mem[sp].t == GM_T_INT);
			mem[sp].i = !(mem[sp].i != 0);
			break;
		case GM_OP_GTR:
		case GM_OP_GEQ:
		case GM_OP_EQU:
		case GM_OP_NEQ:
		case GM_OP_LEQ:
		case GM_OP_LES: {
			SC_NEED(sp >= 1);
			const ScVal b = mem[sp];
			sp--;
			ScVal &a = mem[sp];
			int c; // sign of a - b
			if ((a.t == GM_T_INT || a.t == GM_T_FLOAT) && (b.t == GM_T_INT || b.t == GM_T_FLOAT)) {
				if (a.t == GM_T_INT && b.t == GM_T_INT)
					c = a.i < b.i ? -1 : a.i > b.i;
				else {
					const double x = a.t == GM_T_INT ? (double)a.i : a.d, y = b.t == GM_T_INT ? (double)b.i : b.d;
					SC_NEED(x == x && y == y); // NaN: every comparison false; not worth modelling
					c = x < y ? -1 : x > y;
				}
			} else if (a.t == GM_T_STRING && b.t == GM_T_STRING) {
				c = sc_strcmp(sc, env, a, b);
				SC_NEED(c != 2);
			} else
				return SC_KEEP; // "type mismatch"
			a.t = GM_T_INT;
			a.i = in.op == GM_OP_GTR ? c > 0 : in.op == GM_OP_GEQ ? c >= 0 : in.op == GM_OP_EQU ? c == 0 :
				in.op == GM_OP_NEQ ? c != 0 : in.op == GM_OP_LEQ ? c <= 0 : c < 0;
			break;
		}
		case GM_OP_ADD:
		case GM_OP_SUB:
		case GM_OP_MUL:
		case GM_OP_DIV: {
			// the result keeps the LEFT operand's type (v_tm1->v_value.v_ival += v_top->...v_dval, :2799-2812)
			SC_NEED(sp >= 1);
			const ScVal b = mem[sp];
			sp--;
			ScVal &a = mem[sp];
			SC_NEED((a.t == GM_T_INT || a.t == GM_T_FLOAT) && (b.t == GM_T_INT || b.t == GM_T_FLOAT));
			if (a.t == GM_T_INT && b.t == GM_T_INT) {
				const long long x = a.i, y = b.i;
				long long r;
				if (in.op == GM_OP_DIV) {
					SC_NEED(y != 0 && !(x == -2147483647 - 1 && y == -1));
					r = x / y;
				} else
					r = in.op == GM_OP_ADD ? x + y : in.op == GM_OP_SUB ? x - y : x * y;
				SC_NEED(r >= -2147483647 - 1 && r <= 2147483647); // signed overflow: undefined there
				a.i = (int)r;
			} else {
				const double x = a.t == GM_T_INT ? (double)a.i : a.d, y = b.t == GM_T_INT ? (double)b.i : b.d;
				SC_NEED(in.op != GM_OP_DIV || y != 0.0);
				const double r = in.op == GM_OP_ADD ? x + y : in.op == GM_OP_SUB ? x - y : in.op == GM_OP_MUL ? x * y : x / y;
				if (a.t == GM_T_INT) {
					SC_NEED(r > -2e9 && r < 2e9);
					a.i = (int)r;
				} else
					a.d = r;
			}
			break;
		}
		case GM_OP_MOD: {
			SC_NEED(sp >= 1 && mem[sp].t == GM_T_INT && mem[sp - 1].t == GM_T_INT && mem[sp].i != 0 && mem[sp].i != -1);
			sp--;
			mem[sp].i %= mem[sp + 1].i;
			break;
		}
		case GM_OP_NEG:
			SC_NEED(sp >= 0 && (mem[sp].t == GM_T_INT || mem[sp].t == GM_T_FLOAT));
			if (mem[sp].t == GM_T_INT) {
				SC_NEED(mem[sp].i != -2147483647 - 1);
				mem[sp].i = -mem[sp].i;
			} else
				mem[sp].d = -mem[sp].d;
			break;
		case GM_OP_I_PP:
		case GM_OP_PP_I:
		case GM_OP_I_MM:
		case GM_OP_MM_I: {
			SC_NEED(sp >= 0 && mem[sp].t == GM_T_IDENT);
			const int vi = mem[sp].i;
			SC_NEED(vi >= 0 && vi < sc.n_var && sc.var[vi].special == GM_SV_NONE && var[vi].t == GM_T_INT);
			SC_NEED(var[vi].i > -2000000000 && var[vi].i < 2000000000);
			var[vi].i += (in.op == GM_OP_I_PP || in.op == GM_OP_PP_I) ? 1 : -1;
			mem[sp].t = GM_T_UNKNOWN; // (the slot keeps type T_IDENT there: not a usable value)
			break;
		}
		case GM_OP_FJP:
			SC_NEED(sp >= 0 && mem[sp].t == GM_T_INT);
			if (!mem[sp].i)
				pc = in.a;
			sp = mp = -1;
			break;
		case GM_OP_JMP:
			pc = in.a;
			break;
		case GM_OP_STRF: {
			// do_strf, :2082-2133
			SC_NEED(sp >= 2 && mem[sp].t == GM_T_INT && mem[sp - 1].t == GM_T_INT && mem[sp - 2].t == GM_T_INT);
			int len = mem[sp].i, pos = mem[sp - 1].i;
			const int x = mem[sp - 2].i;
			SC_NEED(x >= 0 && x < sc.n_xel && sc.xel[x].elem >= 0);
			const int d = sc.xel[x].elem, ml = env.len(d);
			if (pos == GM_UNDEF)
				pos = 1;
			else if (pos < 0)
				return SC_KEEP;
			else if (ml == 0)
				pos = 1;
			else if (pos > ml)
				return SC_KEEP;
			pos--;
			SC_NEED(pos >= 0); // (pos = 0 reads one before the element there)
			if (len == 0)
				return SC_KEEP;
			else if (len == GM_UNDEF)
				len = ml - pos;
			else
				len = ml - pos < len ? ml - pos : len;
			SC_NEED(len >= 0);
			sp -= 2;
			mem[sp].t = GM_T_STRING;
			mem[sp].so = env.off(d) + pos;
			mem[sp].sl = len;
			esp--;
			break;
		}
		case GM_OP_SCL:
			switch (in.a) {
			case GM_SC_STRID: {
				// do_scl :1151-1162, strid :1372-1419
				SC_NEED(sp >= 1 && mp >= 0 && mp <= sp && mem[sp - 1].t == GM_T_INT && mem[mp].t == GM_T_INT);
				const ScVal id = mem[sp];
				const int stype = mem[sp - 1].i;
				int idx = -1;
				if (id.t == GM_T_INT) {
					SC_NEED(id.i >= 1 && id.i <= sc.n_xel);
					idx = id.i - 1;
					SC_NEED(stype == sc.sym_se || sc.xel[idx].sym == stype);
				} else if (id.t == GM_T_STRING) {
					for (int s = 0; s < sc.n_xel && idx < 0; s++) {
						if (sc.xel[s].tag < 0)
							continue;
						ScVal tg;
						tg.t = GM_T_STRING;
						tg.so = -sc.xel[s].tag - 1;
						tg.sl = 0;
						for (const char *p = sc.str + sc.xel[s].tag; *p; p++)
							tg.sl++;
						const int c = sc_strcmp(sc, env, tg, id);
						SC_NEED(c != 2);
						if (c == 0 && (sc.xel[s].sym == stype || (sc.xel[s].sym == sc.sym_ss && stype == sc.sym_se)))
							idx = s;
					}
					SC_NEED(idx >= 0);
				} else
					return SC_KEEP;
				sp = mp;
				mp = mem[mp].i;
				mem[sp].t = GM_T_INT;
				mem[sp].i = idx;
				SC_NEED(esp + 1 < 20);
				estk[++esp] = idx;
				break;
			}
			case GM_SC_LENGTH: {
				SC_NEED(sp >= 0 && mp >= 0 && mp <= sp && mem[sp].t == GM_T_STRING && mem[mp].t == GM_T_INT);
				const int len = mem[sp].sl;
				sp = mp;
				mp = mem[mp].i;
				mem[sp].t = GM_T_INT;
				mem[sp].i = len;
				break;
			}
			case GM_SC_MISMATCHES_1:
			case GM_SC_MISPAIRS: {
				SC_NEED(sp >= 2 && mp >= 0 && mp <= sp && mem[sp - 2].t == GM_T_INT && mem[mp].t == GM_T_INT);
				const int x = mem[sp - 2].i;
				SC_NEED(x >= 0 && x < sc.n_xel && sc.xel[x].elem >= 0);
				const int d = sc.xel[x].elem;
				sp = mp;
				mp = mem[mp].i;
				mem[sp].t = GM_T_INT;
				mem[sp].i = in.a == GM_SC_MISPAIRS ? env.mpr(d) : env.mm(d);
				break;
			}
			case GM_SC_PAIRED: {
				// do_scl :1283-1317, paired :1421-1444 (duplexes; triples and quadruples are kept)
				SC_NEED(sp >= 2 && mp >= 0 && mp <= sp && mem[sp].t == GM_T_INT && mem[sp - 1].t == GM_T_INT &&
					mem[sp - 2].t == GM_T_INT && mem[mp].t == GM_T_INT);
				const int x = mem[sp - 2].i;
				SC_NEED(x >= 0 && x < sc.n_xel && sc.xel[x].elem >= 0);
				const int d = sc.xel[x].elem;
				const gm_elem_t &e = env.elem(d);
				const int ml = env.len(d);
				int pos = mem[sp - 1].i, len = mem[sp].i;
				SC_NEED(e.n_mates == 1 && pos >= 1 && pos <= ml && len != 0);
				pos--;
				len = len < 0 ? ml - pos : (ml - pos < len ? ml - pos : len);
				const int d1 = d < e.mates[0] ? d : e.mates[0];
				const int d2 = env.elem(d1).mates[0];
				const int p1 = env.off(d1), p2 = env.off(d2) + ml - 1;
				const unsigned dup = env.pairset(env.elem(d1).pairset).duplex;
				int rv = 1;
				for (int k = 0; k < len && rv; k++) {
					const int c1 = env.ch(p1 + pos + k), c2 = env.ch(p2 - pos - k);
					SC_NEED(c1 >= 0 && c2 >= 0);
					const int b1 = c1 == 'a' ? 0 : c1 == 'c' ? 1 : c1 == 'g' ? 2 : (c1 == 't' || c1 == 'u') ? 3 : 4;
					const int b2 = c2 == 'a' ? 0 : c2 == 'c' ? 1 : c2 == 'g' ? 2 : (c2 == 't' || c2 == 'u') ? 3 : 4;
					rv = (dup >> (b1 * 5 + b2)) & 1u;
				}
				sp = mp;
				mp = mem[mp].i;
				mem[sp].t = GM_T_INT;
				mem[sp].i = rv;
				break;
			}
			case GM_SC_SUBSTR: {
				SC_NEED(sp >= 2 && mp >= 0 && mp <= sp && mem[sp].t == GM_T_INT && mem[sp - 1].t == GM_T_INT &&
					mem[sp - 2].t == GM_T_STRING && mem[mp].t == GM_T_INT);
				const ScVal str = mem[sp - 2];
				const int pos = mem[sp - 1].i;
				int len = mem[sp].i;
				SC_NEED(pos >= 1 && pos <= str.sl && len >= 1);
				len = str.sl - pos + 1 < len ? str.sl - pos + 1 : len;
				sp = mp;
				mp = mem[mp].i;
				mem[sp].t = GM_T_STRING;
				mem[sp].so = str.so >= 0 ? str.so + pos - 1 : str.so - (pos - 1);
				mem[sp].sl = len;
				break;
			}
			default:
				return SC_KEEP; // efn, efn2, bits, sprintf, loc, mismatches(string, pattern)
			}
			break;
		default:
			return SC_KEEP; // acpt, hold, rlse, halt, fcl, mat, ins
		}
	}
#undef SC_NEED
	return SC_KEEP;
}

} // namespace gm

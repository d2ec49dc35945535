/*
 * rm_score_flat.c -- reference-side glue: the compiled MAIN score program
 * (src/score.c: progs[P_MAIN], static there) as a POD image for the device's
 * pre-screen (include/gpumotif_score.h), plus the checks that make skipping a
 * rejected candidate's replay sound.
 *
 * Like rm_replay.c this translation unit #includes the reference's source from
 * where it lies (-DREF_SCORE_C) to reach its file-scope tables; it then stands in
 * for score.o in the link (every external symbol of score.c is defined here, so
 * the archive member is never pulled).  Nothing of the reference is modified.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpumotif_plan.h"
#include "gpumotif_score.h"

#include REF_SCORE_C

static int flat_fail(gm_score_t *sc, const char *why)
{
	sc->present = 0;
	snprintf(sc->why, sizeof sc->why, "%s", why);
	return 0;
}

static int pool_str(gm_score_t *sc, const char *s)
{
	const int n = (int)strlen(s) + 1;
	int o;
	for (o = 0; o < sc->n_str; o += (int)strlen(sc->str + o) + 1)
		if (!strcmp(sc->str + o, s))
			return o;
	if (sc->n_str + n > GM_SC_MAX_STR)
		return -1;
	o = sc->n_str;
	memcpy(sc->str + o, s, (size_t)n);
	sc->n_str += n;
	return o;
}

static int var_of(gm_score_t *sc, IDENT_T **ids, IDENT_T *idp)
{
	int v;
	gm_sc_var_t *gv;
	for (v = 0; v < sc->n_var; v++)
		if (ids[v] == idp)
			return v;
	if (sc->n_var >= GM_SC_MAX_VAR)
		return -1;
	v = sc->n_var++;
	ids[v] = idp;
	gv = &sc->var[v];
	memset(gv, 0, sizeof *gv);
	gv->type = idp->i_type;
	gv->special = !strcmp(idp->i_name, "COMP") ? GM_SV_COMP : !strcmp(idp->i_name, "POS") ? GM_SV_POS :
		!strcmp(idp->i_name, "LEN") ? GM_SV_LEN : !strcmp(idp->i_name, "SLEN") ? GM_SV_SLEN :
		!strcmp(idp->i_name, "NAME") ? GM_SV_NAME : GM_SV_NONE;
	switch (idp->i_type) {
	case T_UNDEF:
		break;
	case T_INT:
		gv->ival = idp->i_val.v_value.v_ival;
		break;
	case T_FLOAT:
		gv->dval = idp->i_val.v_value.v_dval;
		break;
	case T_STRING:
		gv->ival = idp->i_val.v_value.v_pval != NULL ? pool_str(sc, idp->i_val.v_value.v_pval) : -1;
		if (gv->ival < 0)
			gv->type = GM_T_UNKNOWN;
		break;
	default:
		gv->type = GM_T_UNKNOWN;
		break;
	}
	return v;
}

/* result type of a builtin (do_scl, src/score.c:1138-1370) */
static int scl_type(int code)
{
	switch (code) {
	case SC_BITS: case SC_EFN: case SC_EFN2:
		return T_FLOAT;
	case SC_SPRINTF: case SC_SUBSTR:
		return T_STRING;
	default:
		return T_INT;
	}
}

/*
 * elem_of[x]: index in the flattened plan of rm_xdescr[x] (the plan's elements are
 * rm_descr[] in order), -1 / -2 for the contexts.  Call after SE_link /
 * RM_linkscore and after the BEGIN program has run (the variables' values at that
 * point are what MAIN starts from).  Always returns 0; sc->present tells whether the
 * device may pre-screen.
 */
int gm_flatten_score(gm_score_t *sc)
{
	static IDENT_T *ids[GM_SC_MAX_VAR];
	static int sto_var[PROG_SIZE];   /* variable a `sto` writes, else -1 */
	static int rd_var[PROG_SIZE];    /* variable an instruction reads, else -1 */
	static unsigned char da[PROG_SIZE][(GM_SC_MAX_VAR + 7) / 8];
	const INST_T *pm = progs[P_MAIN];
	const int n = l_progs[P_MAIN];
	int vtype[GM_SC_MAX_VAR];
	int pc, x, v, changed;

	memset(sc, 0, sizeof *sc);
	sc->sym_se = SYM_SE;
	sc->sym_ss = SYM_SS;
	for (pc = 0; pc < n; pc++)
		if (pm[pc].i_op == OP_HOLD || pm[pc].i_op == OP_RLSE)
			sc->has_hold = 1;
	if (n <= 0)
		return flat_fail(sc, "no score program");
	if (l_progs[P_END] > 0)
		return flat_fail(sc, "the descriptor has an END section (it may read what MAIN left behind)");
	if (n > GM_SC_MAX_INST)
		return flat_fail(sc, "program too long for the device tables");
	if (rm_n_xdescr > GM_SC_MAX_XEL)
		return flat_fail(sc, "too many elements");
	for (x = 0; x < rm_n_xdescr; x++) {
		const STREL_T *stp = rm_xdescr[x];
		sc->xel[x].elem = stp == rm_lctx ? -1 : stp == rm_rctx ? -2 : (int)(stp - rm_descr);
		sc->xel[x].sym = stp->s_type;
		sc->xel[x].tag = -1;
		if (stp->s_tag != NULL && (sc->xel[x].tag = pool_str(sc, stp->s_tag)) < 0)
			return flat_fail(sc, "string pool full");
	}
	sc->n_xel = rm_n_xdescr;

	/* ---- instructions ---- */
	for (pc = 0; pc < n; pc++) {
		const INST_T *ip = &pm[pc];
		gm_sc_inst_t *o = &sc->inst[pc];
		o->op = (uint8_t)ip->i_op;
		sto_var[pc] = rd_var[pc] = -1;
		switch (ip->i_op) {
		case OP_HOLD: case OP_RLSE:
			return flat_fail(sc, "the program HOLDs / RELEASEs candidates (state carried between candidates)");
		case OP_LDA: case OP_LOD:
			if ((v = var_of(sc, ids, ip->i_val.v_value.v_pval)) < 0)
				return flat_fail(sc, "too many variables");
			o->a = v;
			if (ip->i_op == OP_LOD)
				rd_var[pc] = v;
			break;
		case OP_LDC:
			o->vtype = (uint8_t)ip->i_val.v_type;
			if (ip->i_val.v_type == T_INT)
				o->a = ip->i_val.v_value.v_ival;
			else if (ip->i_val.v_type == T_FLOAT) {
				if (sc->n_dbl >= GM_SC_MAX_DBL)
					return flat_fail(sc, "too many float constants");
				sc->dbl[sc->n_dbl] = ip->i_val.v_value.v_dval;
				o->a = sc->n_dbl++;
			} else if (ip->i_val.v_type == T_STRING) {
				if ((o->a = pool_str(sc, ip->i_val.v_value.v_pval)) < 0)
					return flat_fail(sc, "string pool full");
			} else if (ip->i_val.v_type != T_POS)
				o->vtype = GM_T_UNKNOWN; /* pairset literals: `in` is not pre-screened */
			break;
		case OP_FJP: case OP_JMP: case OP_AND: case OP_IOR: case OP_SCL:
			o->a = ip->i_val.v_value.v_ival;
			break;
		default:
			break;
		}
	}
	sc->n_inst = n;

	/* ---- how variables are used: plain assignments `lda v ... sto` and `lda v; incp` only ---- */
	for (v = 0; v < sc->n_var; v++)
		vtype[v] = sc->var[v].type;
	for (pc = 0; pc < n; pc++) {
		int q, n_sto = 0, sto_at = -1;
		if (pm[pc].i_op != OP_LDA)
			continue;
		v = sc->inst[pc].a;
		if (sc->var[v].special != GM_SV_NONE)
			return flat_fail(sc, "the program writes NAME / COMP / POS / LEN / SLEN");
		if (pc + 1 < n && (pm[pc + 1].i_op == OP_I_PP || pm[pc + 1].i_op == OP_PP_I || pm[pc + 1].i_op == OP_I_MM ||
				   pm[pc + 1].i_op == OP_MM_I)) {
			rd_var[pc + 1] = v; /* read-modify-write: must be assigned already, and stays so */
			continue;
		}
		for (q = pc + 1; q < n && pm[q].i_op != OP_CLS; q++) {
			if (pm[q].i_op == OP_LDA || pm[q].i_op == OP_FJP || pm[q].i_op == OP_JMP)
				return flat_fail(sc, "an assignment the pre-screen cannot follow (nested assignment or jump inside it)");
			if (pm[q].i_op == OP_STO) {
				n_sto++;
				sto_at = q;
			}
		}
		if (n_sto != 1 || q >= n || sto_at != q - 1)
			return flat_fail(sc, "a variable is used in a way the pre-screen cannot follow");
		for (q = pc + 1; q < sto_at; q++) /* && / || inside the right-hand side stay inside it */
			if ((pm[q].i_op == OP_AND || pm[q].i_op == OP_IOR) &&
			    (pm[q].i_val.v_value.v_ival <= q || pm[q].i_val.v_value.v_ival > sto_at))
				return flat_fail(sc, "an assignment the pre-screen cannot follow");
		sto_var[sto_at] = v;
		/* type of the right-hand side, by the interpreter's own rules: arithmetic keeps the
		 * left operand's type, comparisons and most builtins give int (abstract run of the
		 * statement over types only) */
		{
			int ts[64], ms[16], tsp = -1, msp = -1, bad = 0;
			for (q = pc + 1; q < sto_at && !bad; q++) {
				const int op = pm[q].i_op;
				switch (op) {
				case OP_MRK:
					if (tsp + 1 >= 64 || msp + 1 >= 16) { bad = 1; break; }
					ts[++tsp] = T_INT;
					ms[++msp] = tsp;
					break;
				case OP_LDC:
					if (tsp + 1 >= 64) { bad = 1; break; }
					ts[++tsp] = pm[q].i_val.v_type == T_POS ? T_INT : pm[q].i_val.v_type;
					break;
				case OP_LOD: {
					const int w = sc->inst[q].a;
					if (tsp + 1 >= 64) { bad = 1; break; }
					ts[++tsp] = sc->var[w].special == GM_SV_NAME ? T_STRING : sc->var[w].special != GM_SV_NONE ? T_INT : vtype[w];
					if (ts[tsp] == T_UNDEF)
						bad = 1; /* read of something that has no type yet at this point of the text */
					break;
				}
				case OP_SCL:
					if (msp < 0) { bad = 1; break; }
					tsp = ms[msp--];
					ts[tsp] = scl_type(pm[q].i_val.v_value.v_ival);
					break;
				case OP_STRF:
					if (tsp < 2) { bad = 1; break; }
					tsp -= 2;
					ts[tsp] = T_STRING;
					break;
				case OP_GTR: case OP_GEQ: case OP_EQU: case OP_NEQ: case OP_LEQ: case OP_LES: case OP_MAT: case OP_INS:
					if (tsp < 1) { bad = 1; break; }
					tsp--;
					ts[tsp] = T_INT;
					break;
				case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: case OP_MOD:
					if (tsp < 1) { bad = 1; break; }
					tsp--; /* the left operand's type stays */
					break;
				case OP_NOT: case OP_NEG: case OP_AND: case OP_IOR: case OP_NOOP:
					break;
				default:
					bad = 1;
					break;
				}
			}
			if (bad || tsp < 0 || (ts[tsp] != T_INT && ts[tsp] != T_FLOAT && ts[tsp] != T_STRING))
				return flat_fail(sc, "the type of an assignment cannot be told in advance");
			if (vtype[v] == T_UNDEF)
				vtype[v] = ts[tsp];
			else if (sc->var[v].type == T_UNDEF && vtype[v] != ts[tsp])
				/* the variable takes the type of whichever assignment runs first, ever */
				return flat_fail(sc, "a variable is assigned values of different types");
		}
	}

	/* ---- every read must follow a write of the same run on every path: forward
	 * "definitely assigned" dataflow over the instructions ---- */
	for (pc = 0; pc < n; pc++)
		memset(da[pc], pc == 0 ? 0x00 : 0xff, sizeof da[pc]);
	do {
		changed = 0;
		for (pc = 0; pc < n; pc++) {
			unsigned char out[(GM_SC_MAX_VAR + 7) / 8];
			int succ[2], ns = 0, k;
			const int op = pm[pc].i_op;
			memcpy(out, da[pc], sizeof out);
			if (sto_var[pc] >= 0)
				out[sto_var[pc] >> 3] |= (unsigned char)(1u << (sto_var[pc] & 7));
			if (op != OP_JMP && op != OP_RJCT && op != OP_ACPT && op != OP_HALT && pc + 1 < n)
				succ[ns++] = pc + 1;
			if (op == OP_JMP || op == OP_FJP || op == OP_AND || op == OP_IOR) {
				const int t = pm[pc].i_val.v_value.v_ival;
				if (t < 0 || t >= n)
					return flat_fail(sc, "jump out of the program");
				succ[ns++] = t;
			}
			for (k = 0; k < ns; k++) {
				size_t b;
				for (b = 0; b < sizeof out; b++) {
					const unsigned char m = da[succ[k]][b] & out[b];
					if (m != da[succ[k]][b]) {
						da[succ[k]][b] = m;
						changed = 1;
					}
				}
			}
		}
	} while (changed);
	for (pc = 0; pc < n; pc++) {
		v = rd_var[pc];
		if (v < 0 || sc->var[v].special != GM_SV_NONE)
			continue;
		if (sc->var[v].type != T_UNDEF) {
			/* has a value before MAIN starts: fine unless MAIN also writes it (then a read
			 * could see an earlier candidate's value) */
			int q, written = 0;
			for (q = 0; q < n; q++)
				written |= sto_var[q] == v || (rd_var[q] == v && pm[q].i_op != OP_LOD);
			if (!written)
				continue;
		}
		if (!((da[pc][v >> 3] >> (v & 7)) & 1))
			return flat_fail(sc, "a variable can be read before it is written in the same run");
	}
	sc->present = 1;
	return 0;
}

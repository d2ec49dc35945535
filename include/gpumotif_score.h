/*
 * gpumotif_score.h -- the MAIN score program of a descriptor (reference:
 * src/score.c, bytecode :84-122, interpreter RM_score :608-790) flattened to a
 * POD image, so that the device can run it on every candidate at the hit sink and
 * drop the candidates it REJECTs.
 *
 * The device only pre-screens: a candidate is dropped when the program, run with
 * the reference's semantics, reaches `rjct` through instructions of the subset
 * below.  Anything else -- an instruction outside the subset (efn, efn2, bits,
 * sprintf, =~, `in`, user functions), a type the subset does not cover, a run-time
 * error the reference would exit on, ACCEPT, falling off the budget of steps --
 * KEEPS the candidate, and the host replays the whole program on it as before
 * (RM_score + print_match, unchanged).  So the printed output and every SCORE come
 * from the reference's own interpreter; the device saves the replay of rejects.
 *
 * That is only sound when skipping a rejected candidate leaves no trace in later
 * runs of the program.  The flattener (rnamotif_b200/host/rm_score_flat.c) therefore
 * refuses (present = 0: every candidate is replayed) programs that HOLD or RELEASE,
 * programs with an END section, and programs in which some variable can be read
 * before it is written in the same run (the value would come from an earlier
 * candidate).
 */
#ifndef GPUMOTIF_SCORE_H
#define GPUMOTIF_SCORE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_SC_MAX_INST 2048
#define GM_SC_MAX_STR  4096
#define GM_SC_MAX_DBL  128
#define GM_SC_MAX_VAR  96
#define GM_SC_MAX_XEL  102   /* rm_n_xdescr: the elements plus explicit contexts */

/* opcodes: the reference's, src/score.c:84-122 (same numbers) */
enum {
	GM_OP_HALT = 0, GM_OP_NOOP, GM_OP_ACPT, GM_OP_HOLD, GM_OP_RJCT, GM_OP_RLSE, GM_OP_MRK, GM_OP_CLS,
	GM_OP_FCL, GM_OP_SCL, GM_OP_STRF, GM_OP_LDA, GM_OP_LOD, GM_OP_LDC, GM_OP_STO, GM_OP_AND, GM_OP_IOR,
	GM_OP_NOT, GM_OP_MAT, GM_OP_INS, GM_OP_GTR, GM_OP_GEQ, GM_OP_EQU, GM_OP_NEQ, GM_OP_LEQ, GM_OP_LES,
	GM_OP_ADD, GM_OP_SUB, GM_OP_MUL, GM_OP_DIV, GM_OP_MOD, GM_OP_NEG, GM_OP_I_PP, GM_OP_PP_I, GM_OP_I_MM,
	GM_OP_MM_I, GM_OP_FJP, GM_OP_JMP, GM_N_OP
};
/* builtins of `scl`, src/score.c:163-176 */
enum {
	GM_SC_STRID = 0, GM_SC_BITS, GM_SC_EFN, GM_SC_EFN2, GM_SC_LENGTH, GM_SC_LOC, GM_SC_MISMATCHES,
	GM_SC_MISMATCHES_1, GM_SC_MISMATCHES_2, GM_SC_MISPAIRS, GM_SC_PAIRED, GM_SC_SPRINTF, GM_SC_SUBSTR
};
/* value types, src/rnamot.h:49-56 (same numbers); GM_T_UNKNOWN: a constant or variable
 * the device cannot represent -- touching it keeps the candidate */
enum { GM_T_UNDEF = 0, GM_T_INT, GM_T_FLOAT, GM_T_STRING, GM_T_PAIRSET, GM_T_POS, GM_T_IDENT, GM_T_HIT, GM_T_UNKNOWN = 15 };
/* variables the driver sets per candidate (src/find_motif.c:373-389) */
enum { GM_SV_NONE = 0, GM_SV_COMP, GM_SV_POS, GM_SV_LEN, GM_SV_SLEN, GM_SV_NAME };

typedef struct gm_sc_inst {
	uint8_t op;
	uint8_t vtype;   /* type of the operand (ldc), else 0 */
	uint16_t pad;
	int32_t a;       /* int constant / jump target / builtin / variable / offset into str[] / index into dbl[] */
} gm_sc_inst_t;

typedef struct gm_sc_var {
	int32_t type;    /* GM_T_* the variable has when MAIN starts (after BEGIN) */
	int32_t special; /* GM_SV_* */
	int32_t ival;    /* GM_T_INT: the value; GM_T_STRING: offset into str[] */
	int32_t pad;
	double dval;
} gm_sc_var_t;

typedef struct gm_sc_xel {
	int32_t elem;    /* element index in the plan, -1 left context, -2 right context */
	int32_t sym;     /* the reference's SYM_ code of its type (strid compares it, :1390) */
	int32_t tag;     /* offset into str[] or -1 */
} gm_sc_xel_t;

typedef struct gm_score {
	int32_t present;      /* 1: the device may pre-screen with this program */
	int32_t n_inst, n_var, n_xel, n_str, n_dbl;
	int32_t sym_se;       /* SYM_SE: "any element" in strid (:1390,1403) */
	int32_t sym_ss;       /* SYM_SS */
	int32_t has_hold;     /* MAIN contains HOLD or RELEASE (set whatever `present` says) */
	int32_t pad;
	char why[96];         /* present = 0: the reason, for the driver's log */
	gm_sc_inst_t inst[GM_SC_MAX_INST];
	gm_sc_var_t var[GM_SC_MAX_VAR];
	gm_sc_xel_t xel[GM_SC_MAX_XEL];
	double dbl[GM_SC_MAX_DBL];
	char str[GM_SC_MAX_STR];
} gm_score_t;

#ifdef __cplusplus
}
#endif
#endif

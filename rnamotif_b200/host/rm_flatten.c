/*
 * rm_flatten.c -- reference-side glue: turn the compiled descriptor that
 * rnamotif's front end leaves in its globals after SE_link()
 * (src/compile.c:776-820) into the pointer-free gm_plan_t that libgpumotif
 * consumes (include/gpumotif_plan.h).
 *
 * Compiled against the reference headers (rnamot.h, y.tab.h); reads
 * rm_descr[], rm_searches[], rm_sites, rm_lctx/rm_rctx, rm_dminlen/rm_dmaxlen,
 * rm_args and the `windowsize` / `chk_both_strs` builtins.  Writes nothing
 * back.  Anything the device search cannot express is an error here (no CPU
 * fallback): back-references, \( \), \< \>, 8-bit classes, literals outside
 * the IUPAC alphabet, helices longer than GM_MAX_HLEN.
 */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <stdarg.h>

#include "rmdefs.h"
#include "rnamot.h"
#include "y.tab.h"
#include "gpumotif_plan.h"

extern STREL_T rm_descr[];
extern int rm_n_descr;
extern int rm_dminlen, rm_dmaxlen;
extern SEARCH_T **rm_searches;
extern int rm_n_searches;
extern SITE_T *rm_sites;
extern STREL_T *rm_lctx, *rm_rctx;
extern ARGS_T *rm_args;
extern STREL_T *rm_o_stp;
extern char *rm_o_expbuf;

/* regexp bytecodes, src/regexp.c:74-88 */
#define RX_CBRA 2
#define RX_CCHR 4
#define RX_CDOT 8
#define RX_CCL 12
#define RX_CXCL 16
#define RX_CDOL 20
#define RX_CCEOF 22
#define RX_CKET 24
#define RX_CBRC 28
#define RX_CLET 30
#define RX_CBACK 36
#define RX_NCCL 40
#define RX_STAR 01
#define RX_RNGE 03

#define EPS 1e-6 /* src/find_motif.c:15 */

static char *fl_err;
static size_t fl_errlen;

static int fl_fail(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	if (fl_err != NULL && fl_errlen > 0)
		vsnprintf(fl_err, fl_errlen, fmt, ap);
	va_end(ap);
	return -1;
}

/* 4-bit IUPAC code of a lower-case sequence letter; 0 for non-IUPAC letters,
 * -1 for characters that cannot occur in a sequence buffer (FN_fgetseq keeps
 * only isalpha, lower-cased, u->t: src/dbutil.c:105-111) */
static int iupac_code(int ch)
{
	switch (ch) {
	case 'a': return 1;  case 'c': return 2;  case 'g': return 4;
	case 't': return 8;  case 'r': return 5;  case 'y': return 10;
	case 'm': return 3;  case 'k': return 12; case 's': return 6;
	case 'w': return 9;  case 'h': return 11; case 'b': return 14;
	case 'v': return 7;  case 'd': return 13; case 'n': return 15;
	case 'u': return -1;
	}
	if (ch >= 'a' && ch <= 'z')
		return 0;
	return -1;
}

static int type_of(int sym)
{
	switch (sym) {
	case SYM_SS: return GM_SS;
	case SYM_H5: return GM_H5;
	case SYM_H3: return GM_H3;
	case SYM_P5: return GM_P5;
	case SYM_P3: return GM_P3;
	case SYM_T1: return GM_T1;
	case SYM_T2: return GM_T2;
	case SYM_T3: return GM_T3;
	case SYM_Q1: return GM_Q1;
	case SYM_Q2: return GM_Q2;
	case SYM_Q3: return GM_Q3;
	case SYM_Q4: return GM_Q4;
	}
	return -1;
}

static int idx_of(STREL_T *stp)
{
	return stp == NULL ? -1 : (int)(stp - rm_descr);
}

/* class of a CCL/NCCL bitmap over the sequence alphabet */
static int class_of_bitmap(const unsigned char *bm, int neg, uint16_t *cls)
{
	int ch, n_other = 0, n_other_in = 0;
	uint16_t m = 0;
	for (ch = 'a'; ch <= 'z'; ch++) {
		int in = (bm[ch >> 3] >> (ch & 7)) & 1;
		int code = iupac_code(ch);
		if (neg)
			in = !in;
		if (code < 0)
			continue;
		if (code == 0) {
			n_other++;
			n_other_in += in;
		} else if (in)
			m |= (uint16_t)(1u << code);
	}
	if (n_other_in == n_other)
		m |= 1u;
	else if (n_other_in != 0)
		return -1;
	*cls = m;
	return 0;
}

static int add_pos(gm_regex_t *rx, uint16_t cls, int skip, int star, int dot)
{
	int c, p = rx->npos;
	if (p >= GM_RE_MAX_POS)
		return -1;
	for (c = 0; c < 16; c++)
		if ((cls >> c) & 1)
			rx->B[c] |= (uint64_t)1 << p;
	if (skip)
		rx->skip |= (uint64_t)1 << p;
	if (star)
		rx->star |= (uint64_t)1 << p;
	if (dot)
		rx->dot |= (uint64_t)1 << p;
	rx->npos = p + 1;
	return 0;
}

/* s_expbuf -> gm_regex_t.  Bytecode layout: src/regexp.c:158-385. */
static int flatten_expbuf(const char *expbuf, int bol, int mismatch, gm_regex_t *rx, const char *what);

static int flatten_regex(STREL_T *stp, gm_regex_t *rx, const char *what)
{
	return flatten_expbuf(stp->s_expbuf, stp->s_seq[0] == '^', stp->s_mismatch, rx, what);
}

static int flatten_expbuf(const char *expbuf, int bol, int mismatch, gm_regex_t *rx, const char *what)
{
	const unsigned char *ep = (const unsigned char *)expbuf;
	int i, run, maxrun;

	memset(rx, 0, sizeof *rx);
	rx->bol = bol;
	rx->mm_len = -1;
	for (;;) {
		gm_re_item_t it;
		int op = *ep++, base, mod;
		memset(&it, 0, sizeof it);
		if (op == RX_CCEOF)
			break;
		if (op == RX_CDOL) {
			if (*ep != RX_CCEOF)
				return fl_fail("%s: '$' not at end of seq", what);
			rx->eol = 1;
			continue;
		}
		base = op & ~3;
		mod = op & 3;
		if (base == RX_CCHR) {
			int ch = *ep++, code = iupac_code(ch);
			if (code == 0)
				return fl_fail("%s: literal '%c' is outside the IUPAC alphabet", what, ch);
			it.cls = code > 0 ? (uint16_t)(1u << code) : 0;
		} else if (base == RX_CDOT) {
			it.cls = 0xffff;
			it.is_dot = 1;
		} else if (base == RX_CCL || base == RX_NCCL) {
			if (class_of_bitmap(ep, base == RX_NCCL, &it.cls))
				return fl_fail("%s: class distinguishes non-IUPAC letters", what);
			ep += 16;
		} else
			return fl_fail("%s: regex opcode %d (\\( \\) \\< \\> \\n or 8-bit class) not supported on the device", what, op);
		if (mod == RX_STAR)
			it.kind = GM_RE_STAR;
		else if (mod == RX_RNGE) {
			it.kind = GM_RE_RANGE;
			it.lo = *ep++;
			it.hi = *ep++;
		} else if (mod != 0)
			return fl_fail("%s: regex opcode %d not supported", what, op);
		if (rx->eol)
			return fl_fail("%s: '$' not at end of seq", what);
		if (rx->n_items >= GM_RE_MAX_ITEMS)
			return fl_fail("%s: seq too long (> %d items)", what, GM_RE_MAX_ITEMS);
		rx->items[rx->n_items++] = it;
	}

	for (i = 0; i < rx->n_items; i++) {
		gm_re_item_t *it = &rx->items[i];
		int k, rc = 0;
		if (it->kind == GM_RE_ONE)
			rc = add_pos(rx, it->cls, 0, 0, it->is_dot);
		else if (it->kind == GM_RE_STAR)
			rc = add_pos(rx, it->cls, 1, 1, it->is_dot);
		else {
			for (k = 0; k < it->lo && !rc; k++)
				rc = add_pos(rx, it->cls, 0, 0, it->is_dot);
			if (it->hi == 255)
				rc = rc ? rc : add_pos(rx, it->cls, 1, 1, it->is_dot);
			else
				for (k = it->lo; k < it->hi && !rc; k++)
					rc = add_pos(rx, it->cls, 1, 0, it->is_dot);
		}
		if (rc)
			return fl_fail("%s: seq needs more than %d NFA positions", what, GM_RE_MAX_POS);
	}
	for (run = maxrun = 0, i = 0; i < rx->npos; i++) {
		if ((rx->skip >> i) & 1) {
			run++;
			if (run > maxrun)
				maxrun = run;
		} else
			run = 0;
	}
	rx->closure_iters = maxrun;
	if (mismatch > 0) {
		/* mm_seqlen's mmok (src/mm_regexp.c:51-193) already refused
		 * anything that is not fixed-length */
		if (rx->skip != 0)
			return fl_fail("%s: mismatch= on a variable-length seq", what);
		rx->mm_len = rx->npos;
	}
	return 0;
}

static int flatten_pairset(gm_plan_t *pl, PAIRSET_T *ps)
{
	gm_pairset_t g;
	int i, a, b, c, d;

	if (ps == NULL)
		return -1;
	memset(&g, 0, sizeof g);
	g.n_bases = ps->ps_pairs != NULL && ps->ps_n_pairs > 0 ?
		ps->ps_pairs[0].p_n_bases : 2;
	if (ps->ps_mat[0] != NULL) {
		BP_MAT_T *m = (BP_MAT_T *)ps->ps_mat[0];
		for (a = 0; a < N_BCODES; a++)
			for (b = 0; b < N_BCODES; b++)
				if ((*m)[a][b])
					g.duplex |= 1u << (a * 5 + b);
	}
	if (g.n_bases == 3 && ps->ps_mat[1] != NULL) {
		BT_MAT_T *m = (BT_MAT_T *)ps->ps_mat[1];
		for (a = 0; a < N_BCODES; a++)
			for (b = 0; b < N_BCODES; b++)
				for (c = 0; c < N_BCODES; c++)
					if ((*m)[a][b][c]) {
						int k = (a * 5 + b) * 5 + c;
						g.multi[k >> 5] |= 1u << (k & 31);
					}
	} else if (g.n_bases == 4 && ps->ps_mat[1] != NULL) {
		BQ_MAT_T *m = (BQ_MAT_T *)ps->ps_mat[1];
		for (a = 0; a < N_BCODES; a++)
			for (b = 0; b < N_BCODES; b++)
				for (c = 0; c < N_BCODES; c++)
					for (d = 0; d < N_BCODES; d++)
						if ((*m)[a][b][c][d]) {
							int k = ((a * 5 + b) * 5 + c) * 5 + d;
							g.multi[k >> 5] |= 1u << (k & 31);
						}
	}
	/* pairsets are shared by content */
	for (i = 0; i < pl->n_pairsets; i++)
		if (!memcmp(&pl->pairsets[i], &g, sizeof g))
			return i;
	if (pl->n_pairsets >= GM_MAX_PAIRSET)
		return -2;
	pl->pairsets[pl->n_pairsets] = g;
	return pl->n_pairsets++;
}

static int flatten_ctx(gm_plan_t *pl, STREL_T *stp, gm_ctxel_t *cx, const char *what)
{
	memset(cx, 0, sizeof *cx);
	cx->regex = -1;
	if (stp == NULL)
		return 0;
	cx->present = 1;
	cx->minlen = stp->s_minlen;
	cx->maxlen = stp->s_maxlen;
	if (stp->s_seq != NULL) {
		if (stp->s_mismatch > 0)
			return fl_fail("%s: mismatch= on a ctx element is not supported", what);
		if (pl->n_regex >= GM_MAX_REGEX)
			return fl_fail("too many seq= constraints (> %d)", GM_MAX_REGEX);
		if (flatten_regex(stp, &pl->regex[pl->n_regex], what))
			return -1;
		cx->regex = pl->n_regex++;
	}
	return 0;
}

int gm_flatten_plan(gm_plan_t *pl, char *errbuf, size_t errlen)
{
	IDENT_T *ip;
	SITE_T *sip;
	int d, s, k, windowsize;

	fl_err = errbuf;
	fl_errlen = errlen;
	if (errbuf != NULL && errlen > 0)
		errbuf[0] = '\0';

	memset(pl, 0, sizeof *pl);
	pl->magic = GM_PLAN_MAGIC;
	pl->version = GM_PLAN_VERSION;
	if (rm_n_descr <= 0 || rm_n_descr > GM_MAX_DESCR)
		return fl_fail("descriptor has %d elements (1..%d supported)", rm_n_descr, GM_MAX_DESCR);
	pl->n_descr = rm_n_descr;
	pl->n_searches = rm_n_searches;
	pl->dminlen = rm_dminlen;
	pl->dmaxlen = rm_dmaxlen;
	ip = RM_find_id("windowsize");
	if (ip == NULL || ip->i_val.v_value.v_ival <= 0)
		return fl_fail("windowsize undefined or <= 0");
	windowsize = pl->windowsize = ip->i_val.v_value.v_ival;
	pl->strict_helices = rm_args != NULL ? rm_args->a_strict_helices : 0;
	ip = RM_find_id("chk_both_strs");
	pl->chk_both_strs = ip == NULL ? 1 : ip->i_val.v_value.v_ival;

	for (s = 0; s < rm_n_searches; s++) {
		pl->searches[s] = idx_of(rm_searches[s]->s_descr);
		if (s + 1 < rm_n_searches &&
		    rm_searches[s]->s_forward != rm_searches[s + 1]->s_descr)
			return fl_fail("search %d: s_forward is not search %d", s, s + 1);
	}

	for (d = 0; d < rm_n_descr; d++) {
		STREL_T *stp = &rm_descr[d];
		gm_elem_t *e = &pl->elems[d];
		char what[64];
		double pf = stp->s_pairfrac;

		snprintf(what, sizeof what, "element %d (line %d)", d, stp->s_lineno);
		e->type = type_of(stp->s_type);
		if (e->type < 0)
			return fl_fail("%s: unknown element type %d", what, stp->s_type);
		if (stp->s_index != d)
			return fl_fail("%s: s_index %d != position", what, stp->s_index);
		e->searchno = GM_UNDEF;
		for (s = 0; s < rm_n_searches; s++)
			if (rm_searches[s]->s_descr == stp)
				e->searchno = s;
		e->proper = stp->s_attr[SA_PROPER];
		e->ends = stp->s_attr[SA_ENDS];
		e->strict = stp->s_attr[SA_STRICT];
		e->minlen = stp->s_minlen;
		e->maxlen = stp->s_maxlen;
		e->minglen = stp->s_minglen;
		e->maxglen = stp->s_maxglen;
		e->minilen = stp->s_minilen;
		e->maxilen = stp->s_maxilen;
		e->mismatch = stp->s_mismatch;
		e->mispair = stp->s_mispair;
		e->next = idx_of(stp->s_next);
		e->inner = idx_of(stp->s_inner);
		e->outer = idx_of(stp->s_outer);
		e->n_mates = stp->s_n_mates;
		for (k = 0; k < 3; k++)
			e->mates[k] = k < stp->s_n_mates ? idx_of(stp->s_mates[k]) : -1;
		e->scope = stp->s_scope;
		e->n_scopes = stp->s_n_scopes;
		e->scopes = pl->n_scopes;
		if (pl->n_scopes + stp->s_n_scopes > GM_MAX_SCOPES)
			return fl_fail("scope lists too long");
		for (k = 0; k < stp->s_n_scopes; k++)
			pl->scopes[pl->n_scopes++] = idx_of(stp->s_scopes[k]);

		e->pairset = -1;
		if (e->type != GM_SS && stp->s_pairset != NULL) {
			e->pairset = flatten_pairset(pl, stp->s_pairset);
			if (e->pairset == -2)
				return fl_fail("too many distinct pairsets (> %d)", GM_MAX_PAIRSET);
		}

		/* mispair budget: src/find_motif.c:1023-1033 (match_wchlx),
		 * :1122-1130 (match_phlx); same double expression */
		e->pfrac = 0;
		e->mplim = 0;
		e->lentab = -1;
		e->mptab = -1;
		if (e->type != GM_SS) {
			if (stp->s_mispair > 0)
				e->mplim = stp->s_mispair;
			else if (pf < 1.0) {
				e->mplim = (1. - pf) * MIN(stp->s_maxlen, windowsize) + 0.5;
				e->pfrac = 1;
			}
			if (stp->s_maxlen > GM_MAX_HLEN)
				return fl_fail("%s: maxlen %d > %d", what, stp->s_maxlen, GM_MAX_HLEN);
		}
		if (e->type == GM_H5 || e->type == GM_P5 || e->type == GM_T1 || e->type == GM_Q1) {
			/* per-length pairfrac test, src/find_motif.c:1040,1086,1166:
			 * a length hl with mpr mispairs is skipped when
			 * 1.*(hl-mpr)/hl < pairfrac-EPS */
			if (pl->n_lentab + stp->s_maxlen + 1 > GM_LENTAB_SIZE)
				return fl_fail("length tables overflow");
			e->lentab = pl->n_lentab;
			for (k = 0; k <= stp->s_maxlen; k++) {
				int v = 254, m;
				if (e->pfrac && k > 0)
					for (v = 0, m = k; m >= 0; m--)
						if (!(1. * (k - m) / k < pf - EPS)) {
							v = m;
							break;
						}
				pl->lentab[pl->n_lentab++] = (uint8_t)(v > 254 ? 254 : v);
			}
		}
		if (e->type == GM_T1 || e->type == GM_Q2) {
			/* budget by length: src/find_motif.c:1190-1194, 1241-1245 */
			if (pl->n_lentab + stp->s_maxlen + 1 > GM_LENTAB_SIZE)
				return fl_fail("length tables overflow");
			e->mptab = pl->n_lentab;
			for (k = 0; k <= stp->s_maxlen; k++) {
				int v = 0;
				if (stp->s_mispair > 0)
					v = stp->s_mispair;
				else if (pf < 1.0)
					v = (1. - pf) * k + 0.5;
				pl->lentab[pl->n_lentab++] = (uint8_t)(v > 254 ? 254 : v);
			}
		}

		e->regex = -1;
		if (stp->s_seq != NULL) {
			if (pl->n_regex >= GM_MAX_REGEX)
				return fl_fail("too many seq= constraints (> %d)", GM_MAX_REGEX);
			if (flatten_regex(stp, &pl->regex[pl->n_regex], what))
				return -1;
			e->regex = pl->n_regex++;
		}
	}

	for (sip = rm_sites; sip != NULL; sip = sip->s_next) {
		gm_site_t *gs;
		if (pl->n_sites >= GM_MAX_SITES)
			return fl_fail("too many sites (> %d)", GM_MAX_SITES);
		gs = &pl->sites[pl->n_sites++];
		gs->n_pos = sip->s_n_pos;
		if (gs->n_pos < 2 || gs->n_pos > 4)
			return fl_fail("site with %d positions", gs->n_pos);
		gs->pairset = flatten_pairset(pl, sip->s_pairset);
		if (gs->pairset < 0)
			return fl_fail("site pairset: too many distinct pairsets");
		for (k = 0; k < gs->n_pos; k++) {
			POS_T *pp = &sip->s_pos[k];
			gs->pos[k].elem = idx_of((STREL_T *)pp->p_descr);
			gs->pos[k].l2r = pp->p_addr.a_l2r;
			gs->pos[k].offset = pp->p_addr.a_offset;
		}
	}

	/* the literal prefilter, if the front end chose one (-O, default 2.5) */
	memset(&pl->literal, 0, sizeof pl->literal);
	pl->literal.regex = -1;
	if (rm_o_stp != NULL && rm_o_expbuf != NULL && rm_o_stp->s_bestpat.b_lmaxlen != UNBOUNDED &&
	    pl->n_regex < GM_MAX_REGEX) {
		gm_regex_t *rx = &pl->regex[pl->n_regex];
		/* literals are runs of single characters / classes; anything else just
		 * means no prefilter (it is an optimisation, never an error) */
		if (flatten_expbuf(rm_o_expbuf, 0, 1, rx, "best literal") == 0 && rx->npos > 0 && rx->skip == 0 &&
		    !rx->eol) {
			rx->mm_len = rx->npos;
			pl->literal.present = 1;
			pl->literal.regex = pl->n_regex++;
			pl->literal.lmin = rm_o_stp->s_bestpat.b_lminlen;
			pl->literal.lmax = rm_o_stp->s_bestpat.b_lmaxlen;
			pl->literal.mismatch = rm_o_stp->s_mismatch;
		} else if (errbuf != NULL && errlen > 0)
			errbuf[0] = '\0';
	}

	if (flatten_ctx(pl, rm_lctx, &pl->lctx, "left ctx"))
		return -1;
	if (flatten_ctx(pl, rm_rctx, &pl->rctx, "right ctx"))
		return -1;
	return 0;
}

int gm_write_plan(const gm_plan_t *pl, const char *fname)
{
	FILE *fp = fopen(fname, "wb");
	if (fp == NULL)
		return -1;
	if (fwrite(pl, sizeof *pl, 1, fp) != 1) {
		fclose(fp);
		return -1;
	}
	return fclose(fp);
}

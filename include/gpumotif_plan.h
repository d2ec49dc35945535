/*
 * gpumotif_plan.h -- the flattened search plan: a fixed-size POD image of the
 * compiled descriptor that rnamotif's front end leaves in its globals after
 * SE_link() (reference: src/compile.c:776-820).  No pointers, so it can be
 * written to a file, compared byte for byte, and copied into __constant__
 * memory as is.
 *
 * Every field names the reference field it is taken from (STREL_T is
 * src/rnamot.h:228-266, SEARCH_T :268-274, SITE_T/POS_T/ADDR_T :176-194,
 * PAIRSET_T :154-158).  Links that are pointers in the reference are element
 * indices here (-1 = NULL).
 *
 * Sequence alphabet.  The library works on 4-bit IUPAC codes, one per
 * nucleotide: bit0=a bit1=c bit2=g bit3=t, so a=1 c=2 g=4 t=8, r=a|g=5 ...
 * n=15, and 0 = any other letter.  Pairing (src/find_motif.c:1291-1331) sees
 * only the reference's 5 base codes {a,c,g,t,other} (rm_b2bc,
 * src/compile.c:180-187); GM_BCODE(code) gives that.
 */
#ifndef GPUMOTIF_PLAN_H
#define GPUMOTIF_PLAN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_PLAN_MAGIC   0x474d504cu /* "GMPL" */
#define GM_PLAN_VERSION 4

#define GM_UNDEF      (-1)
#define GM_UNBOUNDED  0x7fffffff

#define GM_MAX_DESCR    100  /* RM_DESCR_SIZE, src/compile.c:49 */
#define GM_MAX_REGEX    24
#define GM_MAX_PAIRSET  24
#define GM_MAX_SITES    16
#define GM_MAX_SCOPES   256
#define GM_LENTAB_SIZE  4096
#define GM_RE_MAX_ITEMS 40
#define GM_RE_MAX_POS   63   /* NFA positions; bit npos is "accept" */
#define GM_MAX_HLEN     255  /* longest helix the length tables cover */

/* element types (reference SYM_SS ... SYM_Q4 token values are parser
 * artefacts; these are ours) */
enum {
	GM_SS = 0, GM_H5, GM_H3, GM_P5, GM_P3, GM_T1, GM_T2, GM_T3,
	GM_Q1, GM_Q2, GM_Q3, GM_Q4, GM_N_TYPES
};

/* s_attr[SA_ENDS] bits, src/rnamot.h:210-211 */
#define GM_5PAIRED 1
#define GM_3PAIRED 2
/* s_attr[SA_STRICT] bits, src/rnamot.h:213-214 */
#define GM_5STRICT 1
#define GM_3STRICT 2

/* reference base codes, src/rnamot.h:138-143 */
#define GM_BCODE(code) \
	((code) == 1 ? 0 : (code) == 2 ? 1 : (code) == 4 ? 2 : (code) == 8 ? 3 : 4)

typedef struct gm_elem {
	int32_t type;       /* s_type -> GM_SS ... GM_Q4 */
	int32_t searchno;   /* s_searchno (index into searches[]), -1 if none */
	int32_t proper;     /* s_attr[SA_PROPER] */
	int32_t ends;       /* s_attr[SA_ENDS] */
	int32_t strict;     /* s_attr[SA_STRICT] (0 when strict helices are off) */
	int32_t minlen, maxlen;     /* s_minlen, s_maxlen */
	int32_t minglen, maxglen;   /* s_minglen, s_maxglen */
	int32_t minilen, maxilen;   /* s_minilen, s_maxilen */
	int32_t mismatch;   /* s_mismatch */
	int32_t mispair;    /* s_mispair */
	int32_t pfrac;      /* 1 iff s_mispair <= 0 && s_pairfrac < 1.0
	                       (src/find_motif.c:1023-1033) */
	int32_t mplim;      /* duplex mispair budget of match_wchlx/match_phlx,
	                       src/find_motif.c:1023-1033,1122-1130, evaluated
	                       with the reference's double expression */
	int32_t lentab;     /* H5/P5/T1/Q1: offset into plan.lentab of the
	                       per-length pairfrac table: lentab[hl] = largest
	                       mispair count that passes the test of
	                       src/find_motif.c:1040,1086,1166 at length hl in
	                       [0, maxlen]; 254 = no limit.  -1 otherwise. */
	int32_t mptab;      /* T1 and Q2: offset into plan.lentab of the mispair
	                       budget by length used by match_triplex /
	                       match_4plex (src/find_motif.c:1190-1194,
	                       1241-1245).  -1 otherwise. */
	int32_t next, inner, outer; /* s_next, s_inner, s_outer */
	int32_t n_mates;
	int32_t mates[3];           /* s_mates[] */
	int32_t scope, n_scopes;    /* s_scope, s_n_scopes */
	int32_t scopes;             /* offset into plan.scopes (s_scopes[]) */
	int32_t pairset;            /* index into plan.pairsets or -1 */
	int32_t regex;              /* index into plan.regex or -1 (s_seq == NULL) */
} gm_elem_t;

/* one pairset (PAIRSET_T): the duplex table ps_mat[0] as 25 bits
 * (bit b5*5+b3), and for 3-/4-base pairsets ps_mat[1] as 125 / 625 bits
 * (bit ((b1*5+b2)*5+b3)[*5+b4]). */
typedef struct gm_pairset {
	int32_t  n_bases;     /* 2, 3 or 4 */
	uint32_t duplex;      /* BP_MAT_T ps_mat[0] */
	uint32_t multi[20];   /* BT_MAT_T / BQ_MAT_T ps_mat[1] */
} gm_pairset_t;

/* one regex item = one bytecode of the compiled s_expbuf (src/regexp.c:125-387)
 * restricted to CCHR, CDOT, CCL, NCCL (+STAR, +RNGE); CDOL becomes `eol`. */
enum { GM_RE_ONE = 0, GM_RE_STAR = 1, GM_RE_RANGE = 2 };
typedef struct gm_re_item {
	uint16_t cls;    /* bit c set <=> a nucleotide with IUPAC code c matches */
	uint8_t  kind;   /* GM_RE_ONE / STAR / RANGE */
	uint8_t  is_dot; /* CDOT: never a mismatch in mm_advance */
	uint8_t  lo, hi; /* RANGE \{lo,hi\}; hi == 255: unbounded */
	uint8_t  pad[2];
} gm_re_item_t;

typedef struct gm_regex {
	int32_t n_items;
	int32_t bol;      /* s_seq[0] == '^'  (circf, src/find_motif.c:1818) */
	int32_t eol;      /* trailing CDOL */
	/* bit-parallel NFA derived from items (library side): position i
	 * consumes one nucleotide of class B-membership; `skip` positions may be
	 * bypassed, `star` positions may repeat; bit npos = accept. */
	int32_t  npos;
	int32_t  closure_iters;  /* longest run of consecutive skippable positions */
	int32_t  mm_len;         /* fixed length in mismatch mode, else -1 */
	uint64_t skip, star, dot;
	uint64_t B[16];
	gm_re_item_t items[GM_RE_MAX_ITEMS];
} gm_regex_t;

typedef struct gm_site_pos {
	int32_t elem;    /* p_descr -> element index */
	int32_t l2r;     /* p_addr.a_l2r */
	int32_t offset;  /* p_addr.a_offset */
} gm_site_pos_t;

typedef struct gm_site {
	int32_t n_pos;               /* s_n_pos: 2, 3 or 4 */
	int32_t pairset;             /* s_pairset */
	gm_site_pos_t pos[4];
} gm_site_t;

typedef struct gm_ctxel {
	int32_t present;             /* rm_lctx / rm_rctx != NULL */
	int32_t minlen, maxlen;
	int32_t regex;               /* -1 if no seq= */
} gm_ctxel_t;

/* The reference's literal prefilter (optimize_query, src/compile.c:3315-3392;
 * used by adjust_szero, src/find_motif.c:209-243): the best literal of any
 * seq=, and how far from the start of the motif it can begin.  A start offset
 * can only lead to a candidate if the literal occurs (within its mismatch
 * allowance) at start + d for some d in [lmin, lmax].  Output-neutral. */
typedef struct gm_literal {
	int32_t present;             /* rm_o_stp != NULL */
	int32_t regex;               /* the literal sub-pattern (rm_o_expbuf) as a fixed-length gm_regex_t */
	int32_t lmin, lmax;          /* s_bestpat.b_lminlen / b_lmaxlen of rm_o_stp */
	int32_t mismatch;            /* rm_o_stp->s_mismatch */
	int32_t pad[3];
} gm_literal_t;

typedef struct gm_plan {
	uint32_t magic, version;
	int32_t n_descr;             /* rm_n_descr */
	int32_t n_searches;          /* rm_n_searches */
	int32_t dminlen, dmaxlen;    /* rm_dminlen, rm_dmaxlen */
	int32_t windowsize;          /* builtin `windowsize`, src/find_motif.c:114-128 */
	int32_t strict_helices;      /* rm_args->a_strict_helices */
	int32_t chk_both_strs;       /* builtin, src/rnamot.c:113-117 */
	int32_t n_regex, n_pairsets, n_sites, n_scopes, n_lentab;
	int32_t searches[GM_MAX_DESCR];   /* searches[s]->s_descr as element index;
	                                     s_forward is always searches[s+1]
	                                     (src/compile.c:3290-3296) */
	gm_ctxel_t lctx, rctx;
	gm_literal_t literal;
	gm_elem_t    elems[GM_MAX_DESCR];
	gm_site_t    sites[GM_MAX_SITES];
	gm_pairset_t pairsets[GM_MAX_PAIRSET];
	int32_t      scopes[GM_MAX_SCOPES];
	uint8_t      lentab[GM_LENTAB_SIZE];
	gm_regex_t   regex[GM_MAX_REGEX];
} gm_plan_t;

/* one candidate = one complete assignment that reached the hit sink
 * (src/find_motif.c:362-394): what score.c and print_match read back. */
typedef struct gm_hit_el {
	int32_t off;            /* s_matchoff, 0-based in the searched strand */
	int16_t len;            /* s_matchlen */
	int8_t  n_mispairs;     /* s_n_mispairs */
	int8_t  n_mismatches;   /* s_n_mismatches */
} gm_hit_el_t;

typedef struct gm_hit_hdr {
	uint32_t rec;           /* record number, 0-based, in gm_db_add_record order */
	uint32_t szero;         /* start offset of the window in the searched strand */
	uint32_t seq;           /* rank of this candidate among those of its szero,
	                           in the reference's enumeration order */
	uint8_t  comp;          /* 0 forward, 1 reverse complement */
	uint8_t  pad[3];
	int32_t  lctx_off, lctx_len;   /* set_context results, -1 if no ctx */
	int32_t  rctx_off, rctx_len;
} gm_hit_hdr_t;
/* a hit record is a gm_hit_hdr_t followed by n_descr gm_hit_el_t */

#ifdef __cplusplus
}
#endif
#endif

/*
 * rm_replay.c -- reference-side glue: the tail of rnamotif's hit sink, run on
 * the host over candidates that libgpumotif found.
 *
 * This translation unit #includes the reference's src/find_motif.c from where
 * it lies (path given by -DREF_FIND_MOTIF) so that the reference's own static
 * print_match() and its file-scope state (fm_sid, fm_sbuf, fm_hitbuf ...) are
 * used unchanged.  The CPU search in that file is compiled but never called:
 * RM_find_motif is renamed out of the way and nothing references it.
 *
 * GM_replay_hit() restates the last twenty lines of the sink,
 * src/find_motif.c:373-392: publish NAME/COMP/POS/LEN, run the score program,
 * print unless rejected.
 */
#include <stdio.h>
#include <stdlib.h>

#include "gpumotif_plan.h"

#define RM_find_motif gm_unused_cpu_RM_find_motif
#include REF_FIND_MOTIF
#undef RM_find_motif

void GM_replay_strand(char sid[], char sdef[], int comp, int slen, char sbuf[])
{
	fm_sid = sid;
	fm_sdef = sdef;
	fm_comp = comp;
	fm_slen = slen;
	fm_sbuf = sbuf;
}

/* fp: where print_match writes an accepted hit (stdout, or the driver's capture
 * buffer when it post-filters the output) */
int GM_replay_hit_to(const gm_hit_hdr_t *hdr, const gm_hit_el_t *els, FILE *fp)
{
	IDENT_T *h_idp;
	STREL_T *stp;
	int d, len;

	fm_szero = (int)hdr->szero;
	for (stp = rm_descr, len = 0, d = 0; d < rm_n_descr; d++, stp++) {
		stp->s_matchoff = els[d].off;
		stp->s_matchlen = els[d].len;
		stp->s_n_mispairs = els[d].n_mispairs;
		stp->s_n_mismatches = els[d].n_mismatches;
		len += els[d].len;
	}
	if (rm_lctx != NULL) {
		rm_lctx->s_matchoff = hdr->lctx_off;
		rm_lctx->s_matchlen = hdr->lctx_len;
	}
	if (rm_rctx != NULL) {
		rm_rctx->s_matchoff = hdr->rctx_off;
		rm_rctx->s_matchlen = hdr->rctx_len;
	}
	rm_nval->v_value.v_pval = fm_sid;
	rm_cval->v_value.v_ival = fm_comp;
	rm_pval->v_value.v_ival = fm_comp ? fm_slen - rm_descr[0].s_matchoff
					  : rm_descr[0].s_matchoff + 1;
	rm_lval->v_value.v_ival = len;
	if (RM_score(fm_comp, fm_slen, fm_sbuf, &h_idp) == SA_REJECT)
		return 0;
	print_match(fp, fm_sid, fm_comp, rm_n_descr, rm_descr, h_idp);
	return 1;
}

int GM_replay_hit(const gm_hit_hdr_t *hdr, const gm_hit_el_t *els)
{
	return GM_replay_hit_to(hdr, els, stdout);
}

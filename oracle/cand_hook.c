/*
 * cand_hook.c -- TEST INFRASTRUCTURE.  Instrumented build of the reference
 * search: this wrapper translation unit #includes the reference's
 * src/find_motif.c from where it lies (path given by -DREF_FIND_MOTIF, see
 * oracle/Makefile) and intercepts two calls so that every candidate reaching
 * the hit sink (src/find_motif.c:362-394, i.e. after chk_motif / set_context /
 * chk_sites and before RM_score) is written as one text line to the file named
 * by $GM_CAND_FILE:
 *
 *   rec comp szero action n_descr {off len mispairs mismatches}*n_descr [lctx_off lctx_len rctx_off rctx_len]
 *
 * rec counts records (calls with comp==0), action is RM_score's verdict
 * (0 reject, 1 hold, 2 accept).  Nothing in the reference file is edited.
 */
#include <stdio.h>
#include <stdlib.h>

#define RM_score      gmh_RM_score
#define RM_find_motif gmh_ref_RM_find_motif
#include REF_FIND_MOTIF
#undef RM_score
#undef RM_find_motif

extern int RM_score(int, int, char[], IDENT_T **);

static FILE *gmh_fp;
static int gmh_rec = -1;
static int gmh_init;

static void gmh_open(void)
{
	const char *fn;
	if (gmh_init)
		return;
	gmh_init = 1;
	fn = getenv("GM_CAND_FILE");
	if (fn != NULL && *fn) {
		gmh_fp = fopen(fn, "w");
		if (gmh_fp == NULL) {
			fprintf(stderr, "cand_hook: can't write %s\n", fn);
			exit(1);
		}
	}
}

int gmh_RM_score(int comp, int slen, char sbuf[], IDENT_T **idp)
{
	int rv = RM_score(comp, slen, sbuf, idp);
	gmh_open();
	if (gmh_fp != NULL) {
		int d;
		STREL_T *stp;
		fprintf(gmh_fp, "%d %d %d %d %d", gmh_rec, fm_comp, fm_szero, rv,
			rm_n_descr);
		for (stp = rm_descr, d = 0; d < rm_n_descr; d++, stp++)
			fprintf(gmh_fp, " %d %d %d %d", stp->s_matchoff,
				stp->s_matchlen, stp->s_n_mispairs,
				stp->s_n_mismatches);
		if (rm_lctx != NULL || rm_rctx != NULL)
			fprintf(gmh_fp, " %d %d %d %d",
				rm_lctx ? rm_lctx->s_matchoff : -1,
				rm_lctx ? rm_lctx->s_matchlen : -1,
				rm_rctx ? rm_rctx->s_matchoff : -1,
				rm_rctx ? rm_rctx->s_matchlen : -1);
		fputc('\n', gmh_fp);
	}
	return rv;
}

int RM_find_motif(int n_searches, SEARCH_T *searches[], SITE_T *sites,
	char sid[], char sdef[], int comp, int slen, char sbuf[])
{
	gmh_open(); /* create the file even when nothing is found */
	if (comp == 0)
		gmh_rec++;
	return gmh_ref_RM_find_motif(n_searches, searches, sites, sid, sdef,
		comp, slen, sbuf);
}

/* gm_plan_print -- print a flattened search plan (gpumotif_plan.h) in readable
 * form: elements, searches in order, pseudoknot scopes.  Debug aid.
 *   gcc -I include -o /tmp/gm_plan_print tools/gm_plan_print.c
 *   zcat tests/golden/plans/pk1.plan.gz | /tmp/gm_plan_print */
#include <stdio.h>
#include <stdlib.h>
#include "gpumotif_plan.h"

static const char *tn[] = {"ss","h5","h3","p5","p3","t1","t2","t3","q1","q2","q3","q4"};

int main(void)
{
	static gm_plan_t p;
	if (fread(&p, 1, sizeof p, stdin) != sizeof p) { fprintf(stderr, "short plan\n"); return 1; }
	printf("n_descr %d n_searches %d dmin %d dmax %d window %d strict %d both %d sites %d lit %d (rx %d l %d..%d mm %d)\n",
		p.n_descr, p.n_searches, p.dminlen, p.dmaxlen, p.windowsize, p.strict_helices, p.chk_both_strs, p.n_sites,
		p.literal.present, p.literal.regex, p.literal.lmin, p.literal.lmax, p.literal.mismatch);
	for (int d = 0; d < p.n_descr; d++) {
		const gm_elem_t *e = &p.elems[d];
		printf("el %2d %s s#%2d prop %d ends %d len %d..%d g %d..%d i %d..%d mm %d mpr %d pf %d mplim %d next %d inner %d outer %d mates",
			d, tn[e->type], e->searchno, e->proper, e->ends, e->minlen, e->maxlen, e->minglen, e->maxglen,
			e->minilen, e->maxilen, e->mismatch, e->mispair, e->pfrac, e->mplim, e->next, e->inner, e->outer);
		for (int k = 0; k < e->n_mates; k++) printf(" %d", e->mates[k]);
		printf(" scope %d/%d [", e->scope, e->n_scopes);
		for (int k = 0; k < e->n_scopes; k++) printf(" %d", p.scopes[e->scopes + k]);
		printf(" ] ps %d rx %d", e->pairset, e->regex);
		if (e->regex >= 0) {
			const gm_regex_t *r = &p.regex[e->regex];
			printf(" (bol %d eol %d npos %d mmlen %d items %d)", r->bol, r->eol, r->npos, r->mm_len, r->n_items);
		}
		printf("\n");
	}
	printf("searches:");
	for (int s = 0; s < p.n_searches; s++) printf(" %d", p.searches[s]);
	printf("\n");
	return 0;
}

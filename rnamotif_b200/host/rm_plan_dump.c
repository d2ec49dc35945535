/*
 * rm_plan_dump.c -- compile a descriptor with rnamotif's front end and write
 * the flattened plan (include/gpumotif_plan.h) to a file.  Takes the same
 * arguments as rnamotif plus the environment variable GM_PLAN_OUT=<file>.
 * With GM_SCORE_OUT=<file> also writes the MAIN score program as the device's
 * pre-screen takes it (include/gpumotif_score.h; the BEGIN section is run first,
 * as the driver does).  Used to produce tests/golden/plans/, tests/golden/scores/
 * and by the parity tests.
 */
#include <stdio.h>
#include <stdlib.h>
#include "gpumotif_plan.h"
#include "gpumotif_score.h"
#include "rmdefs.h"
#include "rnamot.h"

extern int gm_rm_compile(int, char *[]);
extern int gm_flatten_plan(gm_plan_t *, char *, size_t);
extern int gm_write_plan(const gm_plan_t *, const char *);
extern int gm_flatten_score(gm_score_t *);

int main(int argc, char *argv[])
{
	static gm_plan_t plan;
	char err[512];
	const char *out = getenv("GM_PLAN_OUT");

	gm_rm_compile(argc, argv);
	if (gm_flatten_plan(&plan, err, sizeof err)) {
		fprintf(stderr, "%s: plan: %s\n", argv[0], err);
		return 2;
	}
	if (out == NULL || !*out) {
		fprintf(stderr, "%s: GM_PLAN_OUT not set\n", argv[0]);
		return 2;
	}
	if (gm_write_plan(&plan, out)) {
		fprintf(stderr, "%s: can't write %s\n", argv[0], out);
		return 2;
	}
	out = getenv("GM_SCORE_OUT");
	if (out != NULL && *out) {
		static gm_score_t score;
		FILE *fp;
		if (RM_fm_init())
			return 2;
		RM_setprog(P_BEGIN);
		RM_score(0, 0, NULL, NULL);
		RM_setprog(P_MAIN);
		gm_flatten_score(&score);
		if ((fp = fopen(out, "wb")) == NULL || fwrite(&score, sizeof score, 1, fp) != 1 || fclose(fp)) {
			fprintf(stderr, "%s: can't write %s\n", argv[0], out);
			return 2;
		}
		fprintf(stderr, "%s: score pre-screen: %s%s%s\n", argv[0], score.present ? "on" : "off (", score.present ? "" : score.why,
			score.present ? "" : ")");
	}
	return 0;
}

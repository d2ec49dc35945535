/*
 * rm_plan_dump.c -- compile a descriptor with rnamotif's front end and write
 * the flattened plan (include/gpumotif_plan.h) to a file.  Takes the same
 * arguments as rnamotif plus the environment variable GM_PLAN_OUT=<file>.
 * Used to produce tests/golden/plans/ and by the parity tests.
 */
#include <stdio.h>
#include <stdlib.h>
#include "gpumotif_plan.h"

extern int gm_rm_compile(int, char *[]);
extern int gm_flatten_plan(gm_plan_t *, char *, size_t);
extern int gm_write_plan(const gm_plan_t *, const char *);

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
	return 0;
}

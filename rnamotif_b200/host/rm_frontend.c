/*
 * rm_frontend.c -- reference-side glue: run rnamotif's own front end
 * (argument parsing, preprocessor, parser, linker, score linker) exactly as
 * the reference driver does before it starts searching
 * (src/rnamot.c:52-111), and stop there.  The compiled descriptor is left in
 * the reference's globals for rm_flatten.c to read.
 */
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include "log.h"
#include "rmdefs.h"
#include "rnamot.h"

extern FILE *yyin;
extern int yyparse(void);
extern int rm_error;
extern ARGS_T *rm_args;
extern int rm_preprocess;
extern int rm_unlink_xdf;
extern STREL_T rm_descr[];
extern int rm_n_descr;
extern int rm_dminlen, rm_dmaxlen;
extern void RM_dump(FILE *, int, int, int, int);

/* returns 0 when the caller should go on to search, otherwise exits the way
 * the reference does (version/-s/-c/-d/-h/-p handling, error exits) */
int gm_rm_compile(int argc, char *argv[])
{
	int early = 0;

	if (RM_init(argc, argv))
		exit(1);
	if (rm_args->a_vopt) {
		fprintf(stderr, "%s: %s.\n", argv[0], VERSION);
		early = 1;
	}
	if (rm_args->a_sopt) {
		RM_dump(stderr, 2, 0, 0, 0);
		early = 1;
	}
	if (early)
		exit(0);

	if (rm_preprocess) {
		rm_args->a_xdfname = RM_preprocessor();
		if (rm_args->a_xdfname == NULL)
			exit(1);
	}
	yyin = fopen(rm_args->a_xdfname, "r");
	if (yyin == NULL) {
		rm_error = TRUE;
		LOG_ERROR("can't read xd-file %s.", rm_args->a_xdfname);
		exit(1);
	}
	if (yyparse()) {
		rm_error = TRUE;
		LOG_ERROR("syntax error.");
	}
	if (rm_unlink_xdf)
		unlink(rm_args->a_xdfname);

	if (!rm_error) {
		if (SE_link(rm_n_descr, rm_descr))
			exit(1);
		RM_linkscore();
		if (rm_args->a_dfname != NULL) {
			fprintf(stderr, "%s: complete descr length: min/max = %d/",
				rm_args->a_dfname, rm_dminlen);
			if (rm_dmaxlen == UNBOUNDED)
				fprintf(stderr, "UNBND\n");
			else
				fprintf(stderr, "%d\n", rm_dmaxlen);
		}
	}
	if (rm_args->a_dopt || rm_args->a_hopt)
		RM_dump(stderr, rm_args->a_dopt, rm_args->a_dopt, rm_args->a_dopt,
			rm_args->a_hopt);
	if (rm_args->a_dopt || rm_args->a_popt)
		RM_dumpscore(stderr);
	if (rm_error)
		exit(1);
	if (rm_args->a_copt)
		exit(0);
	return 0;
}

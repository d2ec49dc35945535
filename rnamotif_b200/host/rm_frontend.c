/*
 * rm_frontend.c -- reference-side glue: run rnamotif's own front end (argument
 * parsing, preprocessor, parser, SE_link, score linker, the -v/-s/-c/-d/-h/-p
 * handling, the data-format check and the opening of the first database file)
 * exactly as the reference driver does before it starts searching, and stop there.
 *
 * Nothing of the reference's main() is restated: this translation unit #includes
 * src/rnamot.c from where it lies (-DREF_RNAMOT_C) with `main` renamed, and with
 * RM_fm_init -- the first thing main() calls once the front end is done
 * (src/rnamot.c:151) -- renamed to a function of ours that jumps back to the
 * caller.  What main() did up to that point is left in the reference's globals
 * (rm_descr[], rm_searches[], rm_args, rm_dbfp = the first database file, opened)
 * for rm_flatten.c / the driver to read; its record loop is never reached.
 * (The real RM_fm_init is compiled in rm_replay.c's translation unit.)
 */
#include <setjmp.h>
#include <stdlib.h>

static jmp_buf gm_front_end_done;

#define main gm_reference_main
#define RM_fm_init gm_front_end_stops_here
#include REF_RNAMOT_C
#undef RM_fm_init
#undef main

int gm_front_end_stops_here(void)
{
	longjmp(gm_front_end_done, 1);
	return 1;
}

/* returns 0 when the caller should go on to search; otherwise the reference's own
 * main() has exited the way it does (usage, -v, -c, errors) */
int gm_rm_compile(int argc, char *argv[])
{
	if (setjmp(gm_front_end_done) == 0) {
		gm_reference_main(argc, argv);
		exit(0); /* (main() does not return: it exits or reaches RM_fm_init) */
	}
	return 0;
}

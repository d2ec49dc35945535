/*
 * frontend_shim.c -- hand-written replacement for the two GENERATED files of
 * the rnamotif front end (y.tab.c from src/rmgrm.y, lex.yy.c from src/rmlex.l).
 *
 * Why it exists: flex and bison are not installed in the build image, and the
 * reference does not vendor its generated parser.  Every other reference
 * source compiles unmodified; the link lacks exactly `yyin` and `yyparse`.
 * This translation unit provides those two symbols.  It contains NO search
 * logic and NO semantic actions of its own: every action calls the reference's
 * own functions (SE_open/SE_addval/SE_close, POS_*, SI_close, PR_*, PARM_add,
 * RM_node, RM_action ... RM_release) in the order the LALR(1) parser generated
 * from src/rmgrm.y:186-545 would call them.
 *
 * Token rules follow src/rmlex.l:40-190.  Grammar follows src/rmgrm.y; the
 * rule each function implements is cited next to it.
 *
 * Built only against the reference headers (rnamot.h, synthesized y.tab.h);
 * see oracle/Makefile.  A site with flex+bison links the generated files
 * instead of this one and nothing else changes.
 */
#include <stdio.h>
#include <ctype.h>
#include <string.h>
#include <stdlib.h>
#include <setjmp.h>

#include "rmdefs.h"
#include "rnamot.h"
#include "y.tab.h"

FILE *yyin;

extern VALUE_T rm_tokval;
extern int rm_context;
extern int rm_lineno;
extern char *rm_wdfname;

extern void RM_hold(NODE_T *);
extern void RM_release(NODE_T *);

#define TOK_EOF 0

typedef struct {
	int sym;
	VALUE_T val;
} TOK_T;

/* ---------------------------------------------------------------- lexer */

static char *lx_buf;
static size_t lx_len, lx_pos;
static int lx_bol; /* at beginning of a line: needed for "^# line" */

#define LA_SIZE 8
static TOK_T la[LA_SIZE];
static int la_head, la_cnt;

static jmp_buf fe_err;

static void fe_fail(void)
{
	longjmp(fe_err, 1);
}

static void *fe_alloc(size_t n)
{
	void *p = malloc(n ? n : 1);
	if (p == NULL) {
		fprintf(stderr, "frontend: out of memory\n");
		exit(1);
	}
	return p;
}

/* file-name interning for "# line N 'file'" (src/rmlex.l:200-262 keeps a tree;
 * a list is enough: names are never freed and compared by content) */
typedef struct fname_s {
	struct fname_s *next;
	char *name;
} FNAME;
static FNAME *fnames;

static char *intern_fname(const char *s, size_t n)
{
	FNAME *f;
	for (f = fnames; f; f = f->next)
		if (strlen(f->name) == n && !strncmp(f->name, s, n))
			return f->name;
	f = fe_alloc(sizeof *f);
	f->name = fe_alloc(n + 1);
	memcpy(f->name, s, n);
	f->name[n] = '\0';
	f->next = fnames;
	fnames = f;
	return f->name;
}

/* src/rmlex.l:200-229 */
static void set_file_info(const char *line, size_t n)
{
	size_t i = 6; /* past "# line" */
	int lnum;
	if (i >= n || !isspace((unsigned char)line[i]))
		return;
	while (i < n && (line[i] == ' ' || line[i] == '\t'))
		i++;
	if (i >= n || !isdigit((unsigned char)line[i]))
		return;
	for (lnum = 0; i < n && isdigit((unsigned char)line[i]); i++)
		lnum = 10 * lnum + line[i] - '0';
	if (i >= n || !isspace((unsigned char)line[i]))
		return;
	while (i < n && isspace((unsigned char)line[i]))
		i++;
	if (i >= n || line[i] != '\'')
		return;
	i++;
	{
		size_t j = i;
		while (j < n && line[j] != '\'')
			j++;
		if (j < n) {
			rm_lineno = lnum;
			rm_wdfname = intern_fname(line + i, j - i);
		}
	}
}

static const struct {
	const char *kw;
	int sym;
} keywords[] = {
	{"parms", SYM_PARMS}, {"descr", SYM_DESCR}, {"sites", SYM_SITES},
	{"score", SYM_SCORE}, {"se", SYM_SE}, {"ctx", SYM_CTX}, {"ss", SYM_SS},
	{"h5", SYM_H5}, {"h3", SYM_H3}, {"p5", SYM_P5}, {"p3", SYM_P3},
	{"t1", SYM_T1}, {"t2", SYM_T2}, {"t3", SYM_T3}, {"q1", SYM_Q1},
	{"q2", SYM_Q2}, {"q3", SYM_Q3}, {"q4", SYM_Q4},
	{"ACCEPT", SYM_ACCEPT}, {"BEGIN", SYM_BEGIN}, {"END", SYM_END},
	{"HOLD", SYM_HOLD}, {"REJECT", SYM_REJECT}, {"RELEASE", SYM_RELEASE},
	{"break", SYM_BREAK}, {"continue", SYM_CONTINUE}, {"else", SYM_ELSE},
	{"for", SYM_FOR}, {"if", SYM_IF}, {"in", SYM_IN}, {"while", SYM_WHILE},
	{NULL, 0}
};

/* length of the FLOAT token at s (src/rmlex.l:98), 0 if none */
static size_t float_len(const char *s, size_t n)
{
	size_t i = 0, nd1 = 0, nd2 = 0, e;
	while (i < n && isdigit((unsigned char)s[i])) { i++; nd1++; }
	if (i < n && s[i] == '.') {
		size_t j = i + 1;
		while (j < n && isdigit((unsigned char)s[j])) { j++; nd2++; }
		if (nd1 == 0 && nd2 == 0)
			return 0;
		i = j;
		/* optional exponent */
		e = i;
		if (e < n && (s[e] == 'e' || s[e] == 'E')) {
			e++;
			if (e < n && (s[e] == '+' || s[e] == '-'))
				e++;
			if (e < n && isdigit((unsigned char)s[e])) {
				while (e < n && isdigit((unsigned char)s[e]))
					e++;
				i = e;
			}
		}
		return i;
	}
	if (nd1 == 0)
		return 0;
	/* digits followed by a mandatory exponent */
	e = i;
	if (e < n && (s[e] == 'e' || s[e] == 'E')) {
		e++;
		if (e < n && (s[e] == '+' || s[e] == '-'))
			e++;
		if (e < n && isdigit((unsigned char)s[e])) {
			while (e < n && isdigit((unsigned char)s[e]))
				e++;
			return e;
		}
	}
	return 0;
}

/* quoted strings, src/rmlex.l:103-138.  The flex rule matches up to (not
 * including) the next quote or newline; if the text so far ends in a backslash
 * it keeps going through the quote (yymore), otherwise one more character is
 * consumed as the closer. */
static char *lex_quoted(int q)
{
	size_t start = lx_pos; /* at the opening quote */
	size_t i = start + 1;
	char *sp;
	size_t n;
	for (;;) {
		while (i < lx_len && lx_buf[i] != q && lx_buf[i] != '\n')
			i++;
		if (i < lx_len && lx_buf[i] == q && lx_buf[i - 1] == '\\' ) {
			i++; /* escaped quote stays in the text */
			continue;
		}
		break;
	}
	/* text = buf[start+1 .. i), closer = buf[i] (whatever it is) */
	n = i - (start + 1);
	sp = fe_alloc(n + 1);
	memcpy(sp, lx_buf + start + 1, n);
	sp[n] = '\0';
	lx_pos = (i < lx_len) ? i + 1 : i;
	return sp;
}

static void lex_one(TOK_T *t)
{
	t->sym = TOK_EOF;
	t->val.v_type = T_UNDEF;
	t->val.v_value.v_pval = NULL;

	for (;;) {
		int c;
		if (lx_pos >= lx_len)
			return;
		c = (unsigned char)lx_buf[lx_pos];
		if (c == '#') {
			size_t e = lx_pos;
			while (e < lx_len && lx_buf[e] != '\n')
				e++;
			if (lx_bol && e - lx_pos >= 6 &&
			    !strncmp(lx_buf + lx_pos, "# line", 6))
				set_file_info(lx_buf + lx_pos, e - lx_pos);
			lx_pos = e;
			lx_bol = 0;
			continue;
		}
		if (c == '\n') {
			lx_pos++;
			lx_bol = 1;
			continue;
		}
		if (c == '\r' || c == ' ' || c == '\t' || c == '\f') {
			lx_pos++;
			if (c != '\r')
				lx_bol = 0;
			continue;
		}
		break;
	}
	lx_bol = 0;

	{
		const char *s = lx_buf + lx_pos;
		size_t n = lx_len - lx_pos;
		int c = (unsigned char)s[0];
		int c1 = n > 1 ? (unsigned char)s[1] : 0;

		if (isalpha(c)) {
			size_t i = 1;
			int k;
			char *sp;
			while (i < n && (isalnum((unsigned char)s[i]) || s[i] == '_'))
				i++;
			lx_pos += i;
			for (k = 0; keywords[k].kw; k++)
				if (strlen(keywords[k].kw) == i &&
				    !strncmp(keywords[k].kw, s, i)) {
					t->sym = keywords[k].sym;
					return;
				}
			sp = fe_alloc(i + 1);
			memcpy(sp, s, i);
			sp[i] = '\0';
			t->sym = SYM_IDENT;
			t->val.v_type = T_STRING;
			t->val.v_value.v_pval = sp;
			return;
		}
		if (isdigit(c) || (c == '.' && isdigit(c1))) {
			size_t fl = float_len(s, n), il = 0;
			char tmp[64];
			while (il < n && isdigit((unsigned char)s[il]))
				il++;
			if (fl > il) {
				size_t m = fl < sizeof tmp - 1 ? fl : sizeof tmp - 1;
				memcpy(tmp, s, m);
				tmp[m] = '\0';
				t->sym = SYM_FLOAT;
				t->val.v_type = T_FLOAT;
				t->val.v_value.v_dval = atof(tmp);
				lx_pos += fl;
			} else {
				size_t m = il < sizeof tmp - 1 ? il : sizeof tmp - 1;
				memcpy(tmp, s, m);
				tmp[m] = '\0';
				t->sym = SYM_INT;
				t->val.v_type = T_INT;
				t->val.v_value.v_ival = atoi(tmp);
				lx_pos += il;
			}
			return;
		}
		if (c == '"') {
			char *sp = lex_quoted('"');
			t->sym = SYM_STRING;
			t->val.v_type = T_STRING;
			t->val.v_value.v_pval = RM_str2seq(sp);
			return;
		}
		if (c == '\'') {
			char *sp = lex_quoted('\'');
			t->sym = SYM_STRING;
			t->val.v_type = T_STRING;
			t->val.v_value.v_pval = sp;
			return;
		}
		if (c == '$') {
			POS_T *posp = fe_alloc(sizeof *posp);
			posp->p_type = SYM_DOLLAR;
			posp->p_lineno = rm_lineno;
			posp->p_tag = NULL;
			posp->p_descr = NULL;
			posp->p_addr.a_l2r = 0;
			posp->p_addr.a_offset = 0;
			t->sym = SYM_DOLLAR;
			t->val.v_type = T_POS;
			t->val.v_value.v_pval = posp;
			lx_pos++;
			return;
		}
#define TWO(a, b, S) if (c == (a) && c1 == (b)) { t->sym = (S); lx_pos += 2; return; }
#define ONE(a, S)    if (c == (a)) { t->sym = (S); lx_pos += 1; return; }
		TWO('&', '&', SYM_AND)
		TWO('=', '=', SYM_EQUAL)
		TWO('=', '~', SYM_MATCH)
		ONE('=', SYM_ASSIGN)
		TWO('!', '~', SYM_DONT_MATCH)
		TWO('!', '=', SYM_NOT_EQUAL)
		ONE('!', SYM_NOT)
		TWO('>', '=', SYM_GREATER_EQUAL)
		ONE('>', SYM_GREATER)
		TWO('<', '=', SYM_LESS_EQUAL)
		ONE('<', SYM_LESS)
		TWO('-', '=', SYM_MINUS_ASSIGN)
		TWO('-', '-', SYM_MINUS_MINUS)
		ONE('-', SYM_MINUS)
		TWO('|', '|', SYM_OR)
		TWO('%', '=', SYM_PERCENT_ASSIGN)
		ONE('%', SYM_PERCENT)
		TWO('+', '=', SYM_PLUS_ASSIGN)
		TWO('+', '+', SYM_PLUS_PLUS)
		ONE('+', SYM_PLUS)
		TWO('*', '=', SYM_STAR_ASSIGN)
		ONE('*', SYM_STAR)
		TWO('/', '=', SYM_SLASH_ASSIGN)
		ONE('/', SYM_SLASH)
		ONE('(', SYM_LPAREN)
		ONE(')', SYM_RPAREN)
		ONE('[', SYM_LBRACK)
		ONE(']', SYM_RBRACK)
		ONE('{', SYM_LCURLY)
		ONE('}', SYM_RCURLY)
		ONE(',', SYM_COMMA)
		ONE(':', SYM_COLON)
		ONE(';', SYM_SEMICOLON)
#undef TWO
#undef ONE
		t->sym = SYM_ERROR;
		lx_pos++;
	}
}

static TOK_T *peek(int k)
{
	if (k >= LA_SIZE)
		fe_fail();
	while (la_cnt <= k) {
		lex_one(&la[(la_head + la_cnt) % LA_SIZE]);
		la_cnt++;
	}
	return &la[(la_head + k) % LA_SIZE];
}

static int psym(int k)
{
	return peek(k)->sym;
}

static TOK_T next(void)
{
	TOK_T t = *peek(0);
	la_head = (la_head + 1) % LA_SIZE;
	la_cnt--;
	return t;
}

static void expect(int sym)
{
	if (psym(0) != sym)
		fe_fail();
	(void)next();
}

/* leaf node carrying the token value: the grammar reads the global rm_tokval
 * at reduction time (src/rmgrm.y:353,507-515,541) */
static NODE_T *leaf(int sym)
{
	TOK_T t;
	if (psym(0) != sym)
		fe_fail();
	t = next();
	rm_tokval = t.val;
	return RM_node(sym, &rm_tokval, 0, 0);
}

/* --------------------------------------------------------------- parser */

static int is_strtype(int s)
{
	switch (s) {
	case SYM_SE: case SYM_CTX: case SYM_SS: case SYM_H5: case SYM_H3:
	case SYM_P5: case SYM_P3: case SYM_T1: case SYM_T2: case SYM_T3:
	case SYM_Q1: case SYM_Q2: case SYM_Q3: case SYM_Q4:
		return 1;
	}
	return 0;
}

static int is_asgn_op(int s)
{
	return s == SYM_ASSIGN || s == SYM_MINUS_ASSIGN || s == SYM_PLUS_ASSIGN ||
	       s == SYM_PERCENT_ASSIGN || s == SYM_SLASH_ASSIGN ||
	       s == SYM_STAR_ASSIGN;
}

static int is_incr_op(int s)
{
	return s == SYM_MINUS_MINUS || s == SYM_PLUS_PLUS;
}

static int is_comp_op(int s)
{
	return s == SYM_DONT_MATCH || s == SYM_EQUAL || s == SYM_GREATER ||
	       s == SYM_GREATER_EQUAL || s == SYM_LESS || s == SYM_LESS_EQUAL ||
	       s == SYM_MATCH || s == SYM_NOT_EQUAL;
}

static NODE_T *p_expr(void);
static NODE_T *p_asgn(void);
static void p_stmt(void);

/* does an asgn start here?  (lval asgn_op ...), lval = ident | auto_lval */
static int at_asgn(void)
{
	if (psym(0) == SYM_IDENT) {
		if (is_asgn_op(psym(1)))
			return 1;
		if (is_incr_op(psym(1)) && is_asgn_op(psym(2)))
			return 1;
		return 0;
	}
	if (is_incr_op(psym(0)) && psym(1) == SYM_IDENT && is_asgn_op(psym(2)))
		return 1;
	return 0;
}

/* does a bare auto_lval start here?  (src/rmgrm.y:504-506) */
static int at_auto_lval(void)
{
	if (is_incr_op(psym(0)) && psym(1) == SYM_IDENT)
		return 1;
	if (psym(0) == SYM_IDENT && is_incr_op(psym(1)))
		return 1;
	return 0;
}

/* lval : ident | auto_lval   (src/rmgrm.y:501-506) */
static NODE_T *p_lval(void)
{
	if (is_incr_op(psym(0))) {
		int op = next().sym;
		NODE_T *id = leaf(SYM_IDENT);
		return RM_node(op, 0, 0, id);
	} else {
		NODE_T *id = leaf(SYM_IDENT);
		if (is_incr_op(psym(0))) {
			int op = next().sym;
			return RM_node(op, 0, id, 0);
		}
		return id;
	}
}

/* pairset : '{' s_list '}'   (src/rmgrm.y:533-540; s_list is right-recursive,
 * so PR_add runs for the LAST string first) */
static NODE_T *p_pairset(void)
{
	NODE_T *strs[64];
	int n = 0, i;
	expect(SYM_LCURLY);
	PR_open();
	for (;;) {
		if (n >= 64)
			fe_fail();
		strs[n++] = leaf(SYM_STRING);
		if (psym(0) == SYM_COMMA) {
			(void)next();
			continue;
		}
		break;
	}
	for (i = n - 1; i >= 0; i--)
		PR_add(strs[i]);
	expect(SYM_RCURLY);
	return PR_close();
}

/* e_list : expr | expr ',' e_list   (src/rmgrm.y:521-524) */
static NODE_T *p_e_list(void)
{
	NODE_T *e = p_expr();
	if (psym(0) == SYM_COMMA) {
		NODE_T *rest;
		(void)next();
		rest = p_e_list();
		return RM_node(SYM_LIST, 0, e, rest);
	}
	return RM_node(SYM_LIST, 0, e, 0);
}

/* a_list : asgn | asgn ',' a_list   (src/rmgrm.y:525-532) */
static NODE_T *p_a_list(void)
{
	NODE_T *a = p_asgn();
	NODE_T *rest = NULL;
	if (psym(0) == SYM_COMMA) {
		(void)next();
		rest = p_a_list();
	}
	if (rm_context == CTX_SCORE)
		return RM_node(SYM_LIST, 0, a, rest);
	return NULL;
}

/* stref : strhdr '(' a_list ')' | strhdr '[' e_list ']'
 * (src/rmgrm.y:226-232,486-500).  In descr/sites context this is also where
 * SE_open/POS_open ... SE_close/POS_close fire. */
static NODE_T *p_stref(int allow_bare)
{
	int type;
	NODE_T *hdr = NULL;
	if (!is_strtype(psym(0)))
		fe_fail();
	type = next().sym;
	if (rm_context == CTX_DESCR)
		SE_open(type);
	else if (rm_context == CTX_SITES)
		POS_open(type);
	else
		hdr = RM_node(type, 0, 0, 0);

	if (psym(0) == SYM_LPAREN) {
		NODE_T *al;
		(void)next();
		al = p_a_list();
		expect(SYM_RPAREN);
		if (rm_context == CTX_DESCR)
			SE_close();
		else if (rm_context == CTX_SITES)
			POS_close();
		else if (rm_context == CTX_SCORE)
			return RM_node(SYM_KW_STREF, 0, hdr, al);
		return NULL;
	}
	if (psym(0) == SYM_LBRACK) {
		NODE_T *el;
		(void)next();
		el = p_e_list();
		expect(SYM_RBRACK);
		return RM_node(SYM_IX_STREF, 0, hdr, el);
	}
	if (!allow_bare)
		fe_fail();
	/* strel : strhdr   (src/rmgrm.y:219-223) */
	if (rm_context == CTX_DESCR)
		SE_close();
	else if (rm_context == CTX_SITES)
		POS_close();
	return hdr;
}

/* primary (src/rmgrm.y:477-482) with literal (:507-513) and fcall (:483-485) */
static NODE_T *p_primary(void)
{
	switch (psym(0)) {
	case SYM_INT:
		return leaf(SYM_INT);
	case SYM_FLOAT:
		return leaf(SYM_FLOAT);
	case SYM_DOLLAR:
		return leaf(SYM_DOLLAR);
	case SYM_STRING:
		return leaf(SYM_STRING);
	case SYM_LCURLY:
		return p_pairset();
	case SYM_LPAREN: {
		NODE_T *e;
		(void)next();
		e = p_expr();
		expect(SYM_RPAREN);
		return e;
	}
	case SYM_IDENT:
		if (psym(1) == SYM_LPAREN) {
			NODE_T *id = leaf(SYM_IDENT), *el;
			expect(SYM_LPAREN);
			el = p_e_list();
			expect(SYM_RPAREN);
			return RM_node(SYM_CALL, 0, id, el);
		}
		return p_lval();
	case SYM_MINUS_MINUS:
	case SYM_PLUS_PLUS:
		return p_lval();
	}
	fe_fail();
	return NULL;
}

/* factor (src/rmgrm.y:452-460) */
static NODE_T *p_factor(void)
{
	if (psym(0) == SYM_MINUS) {
		(void)next();
		return RM_node(SYM_NEGATE, 0, 0, p_primary());
	}
	if (psym(0) == SYM_NOT) {
		(void)next();
		return RM_node(SYM_NOT, 0, 0, p_primary());
	}
	if (is_strtype(psym(0)))
		return p_stref(0);
	return p_primary();
}

/* term : factor | term mul_op factor   (src/rmgrm.y:444-451), left-assoc */
static NODE_T *p_term_from(NODE_T *first)
{
	NODE_T *l = first ? first : p_factor();
	while (psym(0) == SYM_PERCENT || psym(0) == SYM_SLASH || psym(0) == SYM_STAR) {
		int op = next().sym;
		NODE_T *r = p_factor();
		l = RM_node(op, 0, l, r);
	}
	return l;
}

/* a_expr : term | a_expr add_op term   (src/rmgrm.y:437-443), left-assoc */
static NODE_T *p_a_expr_from(NODE_T *first)
{
	NODE_T *l = p_term_from(first);
	while (psym(0) == SYM_PLUS || psym(0) == SYM_MINUS) {
		int op = next().sym;
		NODE_T *r = p_term_from(NULL);
		l = RM_node(op, 0, l, r);
	}
	return l;
}

/* pairing : stref | stref ':' pairing   (src/rmgrm.y:461-468), given the
 * first stref already parsed */
static NODE_T *p_pairing_rest(NODE_T *first)
{
	if (psym(0) == SYM_COLON) {
		NODE_T *s, *rest;
		(void)next();
		s = p_stref(0);
		rest = p_pairing_rest(s);
		if (rm_context == CTX_SCORE)
			return RM_node(SYM_COLON, 0, first, rest);
		return NULL;
	}
	return first;
}

/* compare : site | a_expr | a_expr comp_op a_expr   (src/rmgrm.y:420-424) */
static NODE_T *p_compare(void)
{
	NODE_T *l;
	if (is_strtype(psym(0))) {
		NODE_T *s = p_stref(0);
		if (psym(0) == SYM_COLON || psym(0) == SYM_IN) {
			/* site : pairing SYM_IN pairset   (src/rmgrm.y:260-266) */
			NODE_T *pr = p_pairing_rest(s), *ps;
			expect(SYM_IN);
			ps = p_pairset();
			if (rm_context == CTX_SITES) {
				SI_close(ps);
				return NULL;
			}
			return RM_node(SYM_IN, 0, pr, ps);
		}
		l = p_a_expr_from(s);
	} else
		l = p_a_expr_from(NULL);
	if (is_comp_op(psym(0))) {
		int op = next().sym;
		NODE_T *r = p_a_expr_from(NULL);
		return RM_node(op, 0, l, r);
	}
	return l;
}

/* conj : compare | compare '&&' conj   (src/rmgrm.y:416-419), right-assoc */
static NODE_T *p_conj(void)
{
	NODE_T *l = p_compare();
	if (psym(0) == SYM_AND) {
		NODE_T *r;
		(void)next();
		r = p_conj();
		return RM_node(SYM_AND, 0, l, r);
	}
	return l;
}

/* expr : conj | expr '||' conj   (src/rmgrm.y:412-415), left-assoc */
static NODE_T *p_expr(void)
{
	NODE_T *l = p_conj();
	while (psym(0) == SYM_OR) {
		NODE_T *r;
		(void)next();
		r = p_conj();
		l = RM_node(SYM_OR, 0, l, r);
	}
	return l;
}

/* asgn : lval asgn_op asgn | lval asgn_op expr   (src/rmgrm.y:383-399).
 * Inner assignments reduce (and fire PARM_add/SE_addval) first. */
static NODE_T *p_asgn(void)
{
	NODE_T *lv, *rhs, *n;
	int op;
	lv = p_lval();
	if (!is_asgn_op(psym(0)))
		fe_fail();
	op = next().sym;
	if (at_asgn())
		rhs = p_asgn();
	else
		rhs = p_expr();
	n = RM_node(op, 0, lv, rhs);
	if (rm_context == CTX_PARMS)
		PARM_add(n);
	else if (rm_context == CTX_DESCR || rm_context == CTX_SITES)
		SE_addval(n);
	return n;
}

/* loop_level : SYM_INT | <empty>   (src/rmgrm.y:353-355) */
static NODE_T *p_loop_level(void)
{
	if (psym(0) == SYM_INT)
		return leaf(SYM_INT);
	return NULL;
}

static void p_stmt_list_until_rcurly(void)
{
	/* stmt_list : stmt | stmt stmt_list   (>= 1 statement) */
	do {
		p_stmt();
	} while (psym(0) != SYM_RCURLY);
}

/* stmt (src/rmgrm.y:285-381) */
static void p_stmt(void)
{
	NODE_T *n;
	switch (psym(0)) {
	case SYM_ACCEPT:
		(void)next();
		expect(SYM_SEMICOLON);
		RM_accept();
		return;
	case SYM_REJECT:
		(void)next();
		expect(SYM_SEMICOLON);
		RM_reject();
		return;
	case SYM_BREAK:
		(void)next();
		n = p_loop_level();
		expect(SYM_SEMICOLON);
		RM_break(n);
		return;
	case SYM_CONTINUE:
		(void)next();
		n = p_loop_level();
		expect(SYM_SEMICOLON);
		RM_continue(n);
		return;
	case SYM_HOLD:
		(void)next();
		n = leaf(SYM_IDENT);
		expect(SYM_SEMICOLON);
		RM_hold(n);
		return;
	case SYM_RELEASE:
		(void)next();
		n = leaf(SYM_IDENT);
		expect(SYM_SEMICOLON);
		RM_release(n);
		return;
	case SYM_LCURLY:
		(void)next();
		p_stmt_list_until_rcurly();
		expect(SYM_RCURLY);
		return;
	case SYM_SEMICOLON:
		(void)next();
		return;
	case SYM_FOR:
		/* for_hdr / for_ctrl (src/rmgrm.y:361-381) */
		(void)next();
		expect(SYM_LPAREN);
		if (at_asgn())
			n = p_asgn();
		else if (at_auto_lval())
			n = p_lval();
		else
			n = NULL;
		RM_forinit(n);
		expect(SYM_SEMICOLON);
		if (psym(0) == SYM_SEMICOLON)
			n = NULL;
		else if (at_asgn())
			n = p_asgn();
		else
			n = p_expr();
		RM_fortest(n);
		expect(SYM_SEMICOLON);
		if (at_asgn())
			n = p_asgn();
		else if (at_auto_lval())
			n = p_lval();
		else
			n = NULL;
		RM_forincr(n);
		expect(SYM_RPAREN);
		p_stmt();
		RM_endfor();
		return;
	case SYM_IF:
		/* if_hdr: RM_if fires after the expr, before ')' (src/rmgrm.y:356-360) */
		(void)next();
		expect(SYM_LPAREN);
		n = p_expr();
		RM_if(n);
		expect(SYM_RPAREN);
		p_stmt();
		if (psym(0) == SYM_ELSE) {
			(void)next();
			RM_else();
			p_stmt();
			RM_endelse();
		} else
			RM_endif();
		return;
	case SYM_WHILE:
		(void)next();
		expect(SYM_LPAREN);
		n = p_expr();
		RM_while(n);
		expect(SYM_RPAREN);
		p_stmt();
		RM_endwhile();
		return;
	}
	if (at_asgn()) {
		n = p_asgn();
		expect(SYM_SEMICOLON);
		RM_mark();
		RM_expr(0, n);
		RM_clear();
		return;
	}
	if (at_auto_lval()) {
		n = p_lval();
		expect(SYM_SEMICOLON);
		RM_mark();
		RM_expr(0, n);
		RM_clear();
		return;
	}
	if (psym(0) == SYM_IDENT && psym(1) == SYM_LPAREN) {
		n = p_primary(); /* fcall */
		expect(SYM_SEMICOLON);
		RM_expr(0, n);
		RM_clear();
		return;
	}
	fe_fail();
}

/* action : '{' stmt_list '}'   (src/rmgrm.y:279-281) */
static void p_action(void)
{
	expect(SYM_LCURLY);
	p_stmt_list_until_rcurly();
	expect(SYM_RCURLY);
}

/* rule : pattern action | action   (src/rmgrm.y:271-278).  A leading '{' is an
 * action unless a string follows (then it opens a pairset literal). */
static void p_rule(void)
{
	NODE_T *pat;
	if (psym(0) == SYM_LCURLY && psym(1) != SYM_STRING) {
		p_action();
		return;
	}
	if (psym(0) == SYM_BEGIN) {
		(void)next();
		pat = RM_node(SYM_BEGIN, 0, 0, 0);
	} else if (psym(0) == SYM_END) {
		(void)next();
		pat = RM_node(SYM_END, 0, 0, 0);
	} else
		pat = p_expr();
	RM_action(pat);
	p_action();
	RM_endaction();
}

/* program : parm_part descr_part site_part score_part   (src/rmgrm.y:186-209) */
static void p_program(void)
{
	/* parm_part: optional `parms`, then "asgn ;" until `descr` */
	if (psym(0) == SYM_PARMS) {
		(void)next();
		rm_context = CTX_PARMS;
		do {
			(void)p_asgn();
			expect(SYM_SEMICOLON);
		} while (psym(0) != SYM_DESCR);
	} else if (psym(0) != SYM_DESCR) {
		rm_context = CTX_PARMS;
		do {
			(void)p_asgn();
			expect(SYM_SEMICOLON);
		} while (psym(0) != SYM_DESCR);
	}

	expect(SYM_DESCR);
	rm_context = CTX_DESCR;
	if (!is_strtype(psym(0)))
		fe_fail();
	while (is_strtype(psym(0)))
		(void)p_stref(1);

	if (psym(0) == SYM_SITES) {
		(void)next();
		rm_context = CTX_SITES;
		if (!is_strtype(psym(0)))
			fe_fail();
		while (is_strtype(psym(0))) {
			/* kw_site : kw_pairing SYM_IN pairset   (src/rmgrm.y:253-259) */
			NODE_T *ps;
			(void)p_stref(0);
			while (psym(0) == SYM_COLON) {
				(void)next();
				(void)p_stref(0);
			}
			expect(SYM_IN);
			ps = p_pairset();
			SI_close(ps);
		}
	}

	if (psym(0) == SYM_SCORE) {
		(void)next();
		rm_context = CTX_SCORE;
		do {
			p_rule();
		} while (psym(0) != TOK_EOF);
		RM_accept();
	}

	if (psym(0) != TOK_EOF)
		fe_fail();
}

int yyerror(char *msg)
{
	fprintf(stderr, "yyerror: %s\n", msg);
	return 0;
}

int yyparse(void)
{
	size_t cap = 1 << 16, n = 0;
	int c;

	lx_buf = fe_alloc(cap);
	while ((c = getc(yyin)) != EOF) {
		if (n + 1 >= cap) {
			cap *= 2;
			lx_buf = realloc(lx_buf, cap);
			if (lx_buf == NULL) {
				fprintf(stderr, "frontend: out of memory\n");
				exit(1);
			}
		}
		lx_buf[n++] = (char)c;
	}
	lx_buf[n] = '\0';
	lx_len = n;
	lx_pos = 0;
	lx_bol = 1;
	la_head = la_cnt = 0;

	if (setjmp(fe_err)) {
		yyerror("syntax error");
		return 1;
	}
	p_program();
	return 0;
}

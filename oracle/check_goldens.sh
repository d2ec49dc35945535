#!/bin/sh
# oracle/check_goldens.sh -- TEST INFRASTRUCTURE.
# Pins the reference build (oracle/_ref/rnamotif = reference sources +
# frontend shim) against every golden the reference's own `make test` holds
# (test/Makefile:1-245): 12 slack + 12 strict descriptors over
# gbrna.111.0.fastn, each piped through rmfmt -l and diffed with X.chk.
# Usage: oracle/check_goldens.sh [rnamotif-binary]   (default oracle/_ref/rnamotif)
# Prints one line per test: name PASS|FAIL md5(raw stdout) hits
here=$(cd "$(dirname "$0")" && pwd)
bin=${1:-$here/_ref/rnamotif}
case "$bin" in /*) ;; *) bin=$(pwd)/$bin ;; esac
data=$here/_ref/data
EFNDATA=$data/efndata; export EFNDATA
cd "$data/test" || exit 2
fail=0
for t in nanlin pk1 pk_j1+2 qu+tr score.1 score.2 trna mp.ends efn sprintf bulge getbest; do
	for mode in slack strict; do
		if [ $mode = slack ]; then
			name=$t; flags=""
		else
			name=$t.strict; flags="-sh -context -Dctx_maxlen=5"
		fi
		"$bin" $flags -descr $name.descr gbrna.111.0.fastn > /tmp/gm_$$.raw 2>/tmp/gm_$$.err
		"$here/_ref/rmfmt" -l < /tmp/gm_$$.raw > /tmp/gm_$$.fmt 2>/dev/null
		md5=$(md5sum < /tmp/gm_$$.raw | cut -d' ' -f1)
		hits=$(grep -c '^>' /tmp/gm_$$.raw)
		if cmp -s /tmp/gm_$$.fmt $name.chk; then
			echo "$name PASS $md5 $hits"
		else
			echo "$name FAIL $md5 $hits"; fail=1
		fi
	done
done
rm -f /tmp/gm_$$.raw /tmp/gm_$$.err /tmp/gm_$$.fmt
exit $fail

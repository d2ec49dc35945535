"""rnamotif_b200 -- B200-native descriptor search for rnamotif (the hot path of
src/find_motif.c), behind the C ABI of include/gpumotif.h."""

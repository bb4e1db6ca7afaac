# usage: hnsw_exp.sh  -- hnsw_search_kernel variants (gather width G, min resident blocks) at ef = 200 and 50
for cfg in "4 6" "4 8" "2 8" "8 3"; do set -- $cfg
  for ef in 200 50; do echo "G=$1 MINB=$2 ef=$ef: $(NB200_HNSW_G=$1 NB200_HNSW_MINB=$2 timeout 200 python tools/hnsw_prof.py 100000 960 $ef 2>&1 | tail -n 1)"; done; done

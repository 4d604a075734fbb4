# usage: bash profiles/tl_variants.sh NAME...   -- steady-state timing of libfused_NAME.so variants (profiles/build_fused_variant.sh)
for v in "$@"; do echo "== $v"; MIRL_B200_LIB=mirl_b200/_build/variants/libfused_$v.so timeout 100 python profiles/dbg_train_fused.py time 2>&1 | grep -E "B=256: launch of  400|B=1024: launch of  100" | head -2; done

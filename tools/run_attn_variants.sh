# usage: bash tools/run_attn_variants.sh "<variant suffixes>"   (A/B timing of attention kernel builds, B=32)
for v in $1; do
  [ "$v" = base ] && v=""
  VITTF_LIB=vittf_b200/libvittf_b200$v.so B=${B:-32} python tools/attn_time.py
done

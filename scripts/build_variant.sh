# usage: bash scripts/build_variant.sh <name> <extra nvcc flags...>  -> mcts.jl_b200/csrc/variants/lib<name>.so (tuning experiments)
set -e
NAME=$1; shift
D=mcts.jl_b200/csrc
mkdir -p $D/variants/$NAME
make -C $D embedded_sources.inc > /dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off $*"
for f in capi plugin_rt inst_gw_uct_fast inst_gw_dpw_fast inst_gw_uct inst_gw_dpw inst_mc_uct inst_mc_dpw inst_box_dpw; do
  nvcc $FLAGS -c -o $D/variants/$NAME/$f.o $D/$f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $D/variants/lib$NAME.so $D/variants/$NAME/*.o -ldl
rm -rf $D/variants/$NAME
ls -la $D/variants/lib$NAME.so

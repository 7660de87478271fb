# Diagnostics (not a test): builds ab/lib_<name>.so = the current csrc/ with the listed files taken from another
# commit, for same-box A/B runs of whole libraries:  MZB200_LIB=ab/lib_<name>.so python bench.py ...
#   usage: bash tests/ab_build.sh <name> <commit> <file.cu> [...]
set -e
name=$1; commit=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
d=$(mktemp -d)
mkdir -p "$d/pkg/csrc" "$d/include" "$root/ab"
cp "$root"/model-based-rl_b200/csrc/*.cu "$root"/model-based-rl_b200/csrc/*.cuh "$root"/model-based-rl_b200/csrc/*.h \
   "$root"/model-based-rl_b200/csrc/*.inc "$root"/model-based-rl_b200/csrc/Makefile "$d/pkg/csrc/"
cp "$root/include/mzb200.h" "$d/include/"
for f in "$@"; do git -C "$root" show "$commit:model-based-rl_b200/csrc/$f" > "$d/pkg/csrc/$f"; done
make -C "$d/pkg/csrc" -j8 > /dev/null
cp "$d/pkg/libmzb200.so" "$root/ab/lib_$name.so"
rm -rf "$d"
echo "built ab/lib_$name.so"

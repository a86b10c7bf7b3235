for rep in 1 2 3; do
for lib in tools/micro/libvqb200_old.so ""; do
for shape in "1048576 8192 256" "262144 1024 512 f32" "524288 16384 512 f32"; do
VQB_LIB_OVERRIDE=$lib python tools/tc_knobs.py $shape 2>&1 | tail -1
done; done; done

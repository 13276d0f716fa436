cd tools
for sh in 4 82 42; do
  ./slot_bench 0 $sh 1 1025 1 2
  ./slot_bench 0 $sh 1 1025 16 2
  ./slot_bench 0 $sh 1 1025 32 2
  ./slot_bench 0 $sh 1 65536 1 1
done
./slot_bench 0 82 2 1025 16 2
./slot_bench 0 82 2 65536 1 1
for sh in 4 82 42; do
  ./slot_bench 1 $sh 1 1025 16 2
  ./slot_bench 1 $sh 1 65536 1 1
done

cd tools
for sh in 8 16; do
  for w in 1 2 4; do
  ./slot_bench 0 $sh $w 1025 1 2
  ./slot_bench 0 $sh $w 1025 16 2
  done
  ./slot_bench 0 $sh 2 1025 32 2
  ./slot_bench 0 $sh 2 65536 1 1
done
./slot_bench 1 8 2 1025 16 2
./slot_bench 1 8 2 65536 1 1

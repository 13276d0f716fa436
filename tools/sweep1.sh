cd tools
for g in 1 2 4; do for w in 1 2 4; do
  timeout 60 ./slot_bench 0 $g $w 1025 1 2
  timeout 60 ./slot_bench 0 $g $w 1025 16 2
done; done
for g in 1 2 4; do timeout 60 ./slot_bench 0 $g 2 1025 32 2; timeout 60 ./slot_bench 0 $g 2 16400 1 2; timeout 60 ./slot_bench 0 $g 2 65536 1 1; done
for g in 1 2 4; do timeout 60 ./slot_bench 1 $g 2 1025 1 2; timeout 60 ./slot_bench 1 $g 2 1025 16 2; timeout 60 ./slot_bench 1 $g 2 65536 1 1; done

# round 2, lab 8: CTA shapes again, now that every shape starts its tiles longest rays first
mkdir -p gpurun_out
( LAB_COMBOS="4,32,42;4,32,8;4,32,82;4,32,4;4,32,41;4,32,21;4,32,2;4,24,42;3,36,42" timeout 1500 python tools/lab/variants.py ) > gpurun_out/lab_variants5.log 2>&1; echo "rc=$?" >> gpurun_out/lab_variants5.log
cat gpurun_out/lab_variants5.log

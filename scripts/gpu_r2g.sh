#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/pytest_all.log; tail -5 $O/pytest_all.log
timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_2cta.txt 2>&1; grep "enc2\|rec\|dec\|forward" $O/layers_2cta.txt
S3D_IGEMM_ONE_CTA=1 timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_1cta.txt 2>&1; grep "enc2\|rec\|dec\|forward" $O/layers_1cta.txt
S3D_IGEMM_ONE_CTA=1 S3D_IGEMM_TS1=1 timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_1cta_ts1.txt 2>&1; grep "enc2\|rec\|dec\|forward" $O/layers_1cta_ts1.txt

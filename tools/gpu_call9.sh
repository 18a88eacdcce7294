#!/bin/bash
mkdir -p gpurun_out/r02
python tools/svd_drivers.py > gpurun_out/r02/svd_drivers.txt 2>&1; cat gpurun_out/r02/svd_drivers.txt
python tools/profile_itebd.py > gpurun_out/r02/itebd_host_profile.txt 2>&1; head -70 gpurun_out/r02/itebd_host_profile.txt
python tools/itebd_kernels.py > gpurun_out/r02/itebd_kernels.txt 2>&1; head -45 gpurun_out/r02/itebd_kernels.txt | cut -c1-200

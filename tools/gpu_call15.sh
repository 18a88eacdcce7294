for k in 1e-2 1e-3 1e-4 1e-6; do
echo "kappa $k"; TOR10_SVD_ABS=$k timeout 200 python -c "
import bench, json; d=bench.itebd_leg(); print({x: d[x] for x in ('steps_per_s','eager_steps_per_s','max_abs_dE_vs_cpu_port','max_abs_dE_eager_vs_cpu_port','steps')})"
done
TOR10_SVD_ABS=1e-3 timeout 120 python -m pytest tests/test_gpu_svd.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3

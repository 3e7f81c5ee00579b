set -x
python -m pytest tests/test_gpu_prover.py tests/test_r1cs.py tests/test_mpc.py tests/test_linear_proof.py -m gpu -x -q 2>&1 | tail -8
python benchmarks/prover_batch.py > gpurun_out/r2f_prover.json 2> gpurun_out/r2f_prover.err; cat gpurun_out/r2f_prover.json; tail -3 gpurun_out/r2f_prover.err
python benchmarks/r1cs_shuffle.py 32769 --cpu > gpurun_out/r2f_r1cs.json 2> gpurun_out/r2f_r1cs.err; cat gpurun_out/r2f_r1cs.json; tail -3 gpurun_out/r2f_r1cs.err

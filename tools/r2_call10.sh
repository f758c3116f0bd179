# Round-2 1-GPU call 10: rasteriser variants (bit-exact tests + timing), VAE joint dealing test
mkdir -p gpurun_out
ICB_RASTER_STAGE=1 timeout 300 python -m pytest tests/test_gpu_raster.py tests/test_mesh.py tests/test_io_formats.py tests/test_gpu_vae.py -q -m gpu > gpurun_out/c10_tests_stage1.log 2>&1; echo "exit $?" >> gpurun_out/c10_tests_stage1.log
ICB_RASTER_STAGE=0 timeout 300 python -m pytest tests/test_gpu_raster.py -q -m gpu > gpurun_out/c10_tests_stage0.log 2>&1; echo "exit $?" >> gpurun_out/c10_tests_stage0.log
ICB_RASTER_STAGE=0 timeout 300 python tools/raster_sweep.py --no-cpu > gpurun_out/c10_sweep_stage0.log 2>&1; cp gpurun_out/raster_sweep.json gpurun_out/c10_sweep_stage0.json
ICB_RASTER_STAGE=1 timeout 300 python tools/raster_sweep.py --no-cpu > gpurun_out/c10_sweep_stage1.log 2>&1; cp gpurun_out/raster_sweep.json gpurun_out/c10_sweep_stage1.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:raymarch -c 1 -o gpurun_out/r2_raymarch python - > gpurun_out/c10_ncu.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from infinicube_b200.raster import PinholeCamera, VoxelGrid, synthetic as syn
dev = torch.device('cuda:0')
pts, sem, inst, _ = syn.synthetic_scene(256, voxel_size=0.2)
g = VoxelGrid(torch.from_numpy(pts).to(dev), [0.2] * 3, [0.1] * 3, torch.from_numpy(sem).to(dev), torch.from_numpy(inst).to(dev))
cam = PinholeCamera.from_numpy(syn.DEFAULT_INTRINSICS, device=dev)
poses = torch.from_numpy(syn.synthetic_poses(256, n=93, voxel_size=0.2)).to(dev)
cam.render_voxel_buffers(poses, g); torch.cuda.synchronize()
PY
grep -h "passed\|failed\|^exit" gpurun_out/c10_tests_stage1.log gpurun_out/c10_tests_stage0.log | tail -4
grep -h -o '"S": [0-9]*\|"render_ms_93cams": [0-9.]*' gpurun_out/c10_sweep_stage0.log | tr '\n' ' '; echo
grep -h -o '"S": [0-9]*\|"render_ms_93cams": [0-9.]*' gpurun_out/c10_sweep_stage1.log | tr '\n' ' '; echo
ls -la gpurun_out/r2_raymarch.ncu-rep

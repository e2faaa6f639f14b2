mkdir -p gpurun_out/r3l
# layered at 224x224, batch 256, one pass: 26 kernels; capture the 4 heaviest kinds in full (skip the first pass = 26 launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1x1_tcgen05|dwconv3x3_band|maxpool_band|conv_im2col" -s 26 -c 8 -o gpurun_out/r3l/layered224 -f python tools/run_once.py 256 layered 2 224 > gpurun_out/r3l/ncu.log 2>&1
tail -3 gpurun_out/r3l/ncu.log; ls -la gpurun_out/r3l

mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/c1_32 python tools/profile_conv.py c 8 128 1024 32 32 1 > gpurun_out/p1.log 2>&1
$NCU -k regex:conv_tcw_kernel -s 3 -c 1 -f -o gpurun_out/w_32 python tools/profile_conv.py w 8 128 1024 32 32 1 > gpurun_out/p2.log 2>&1
$NCU -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/c1_256s python tools/profile_conv.py c 8 4 32 256 256 1 > gpurun_out/p3.log 2>&1
$NCU -k regex:conv_tcw_kernel -s 3 -c 1 -f -o gpurun_out/w_256s python tools/profile_conv.py w 8 8 64 256 256 1 > gpurun_out/p4.log 2>&1
for a in "c 8 128 1024 32 32 1" "t 8 128 1024 32 32 1" "w 8 128 1024 32 32 1" "c 8 64 512 64 64 1" "w 8 64 512 64 64 1" "c 8 4 32 256 256 1" "w 8 8 64 256 256 1" "c 8 128 1024 32 64 2" "t 8 128 1024 32 64 2" "w 8 128 1024 32 64 2"; do python tools/profile_conv.py $a; done > gpurun_out/conv_times.log 2>&1
tail -2 gpurun_out/p*.log; cat gpurun_out/conv_times.log

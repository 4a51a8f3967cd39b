tools/ncu_full.sh 5 8 "long:k_dwt_.*_long|k_notch_cplx:0:20" 
tools/ncu_full.sh 4 8 "ls:k_lightsheet_final|k_sort_cells|k_window_percentile:0:4"
tools/ncu_full.sh 3 8 "c3:k_gauss5|k_block_reduce|k_epilogue:0:3"

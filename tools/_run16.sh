o=gpurun_out/sweep_blob_rowminor_r02.txt; rm -f $o
python tools/prof_blob.py 4096 3 15 >> $o 2>&1
python tools/prof_blob.py 4096 3 15 r1_row_minor=1 >> $o 2>&1
python tools/prof_blob.py 4096 3 13 >> $o 2>&1
python tools/prof_blob.py 4096 3 13 r1_row_minor=1 >> $o 2>&1
cat $o

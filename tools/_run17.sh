o=gpurun_out/sweep_blob_rowminor_r02b.txt; rm -f $o
python tools/prof_blob.py 4096 3 15 r1_row_minor=1 tree_chunk_rows=4096 >> $o 2>&1
python tools/prof_blob.py 4096 3 15 r1_row_minor=1 tree_chunk_rows=1024 >> $o 2>&1
python tools/prof_blob.py 4096 3 15 r1_row_minor=0 tree_chunk_rows=4096 >> $o 2>&1
cat $o

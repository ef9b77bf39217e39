mkdir -p gpurun_out
( timeout 200 eesen_b200/bin/cluster_exchange2 ) > gpurun_out/r02x_cluster_bulk.txt 2>&1
head -24 gpurun_out/r02x_cluster_bulk.txt

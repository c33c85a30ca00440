for v in "FS2K_PDL=0 VARIANT=all" "FS2K_PDL=1 VARIANT=nosink" "FS2K_PDL=0 VARIANT=nosink"; do
  env $v timeout 300 python profiles/debug_graph_vs_eager.py 2>&1 | grep -v Warning | grep -E "VARIANT|step|graph |eager " | head -14
done

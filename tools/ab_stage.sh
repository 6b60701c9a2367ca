echo "== base (commit feed641)"; PSB_LIB=$PWD/tools/bin/lib_base.so timeout 200 python tools/tags_probe.py 2>&1 | grep "tags="
for st in 0 1; do echo "== new PSB_STAGE=$st"; PSB_STAGE=$st timeout 200 python tools/tags_probe.py 2>&1 | grep "tags="; done

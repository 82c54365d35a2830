for v in "$@"; do if [ "$v" = default ]; then unset F1L_LIB; else export F1L_LIB=$PWD/f1tenth_planning_b200/lib/variants/libf1l_$v.so; fi; echo -n "$v: "; python tools/run_c2.py; done

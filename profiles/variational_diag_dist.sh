for v in default hostfed hostphilox; do for s in 0 1 2 3 4 5 6 7; do DIAG_SEED=$s python profiles/variational_diag.py mhd_p_dynamic_variational 0 12 $v 2>&1 | tail -1 | cut -c1-90; done; done

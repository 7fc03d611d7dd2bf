for l in 8 16; do for w in 1 2 4 8; do echo -n "lanes $l warps/CTA $w: "; B2G_LANES=$l B2G_NARROW_WARPS=$w python tools/prof_cfg.py cfg2 4096 3 2>&1 | grep -v Warn | cut -c1-60; done; done

"""Read the per-CTA epilogue counters a -DLK_EPI_PROF build leaves in the LK_UMMA_DUMP file (bring-up aid)."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64)[: 148 * 10].reshape(148, 10)
a = a[(a[:, 7] > 0) & (a[:, 7] < 1 << 40)]
m = a.mean(axis=0)
units = m[7]
print(f"CTAs {len(a)}  units/CTA {units:.1f}  cycles per unit: total {m[8] / units:.0f} (in-loop {m[4] / units:.0f} + tfull wait {m[0] / units:.0f})")
print(f"  per unit: tmem ld {m[1] / units:.0f}  fast math {m[2] / units:.0f}  select {m[3] / units:.0f};  turns/unit {m[6] / units:.1f}, turns with a hit {100 * m[5] / m[6]:.1f} %")

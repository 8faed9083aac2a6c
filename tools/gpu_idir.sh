python - <<"EOF"
import numpy as np, cv2, os
os.makedirs("/tmp/fr", exist_ok=True)
rng = np.random.default_rng(0)
base = cv2.GaussianBlur(rng.integers(0,256,(270,480,3),dtype=np.uint8),(0,0),5)
for i in range(5):
    cv2.imwrite("/tmp/fr/%03d.png"%i, np.roll(base, 3*i, axis=1))
print("made", len(os.listdir("/tmp/fr")))
EOF
timeout 200 python tools/interpolate_dir.py --input-dir /tmp/fr --output-dir /tmp/out1 --upsample-rate 4 --amp --channels-last 2>&1 | tail -3
ls /tmp/out1 | wc -l
timeout 200 python tools/interpolate_dir.py --input-dir /tmp/fr --output-dir /tmp/out2 --upsample-rate 4 --n-frames 4 --bottleneck CLSTM --amp --channels-last --save-flows 2>&1 | tail -3
ls /tmp/out2 | wc -l

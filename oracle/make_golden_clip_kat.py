"""Known-answer fixtures for CLIP from the reference's own test data (TEST INFRASTRUCTURE; needs /root/reference).

The reference holds two known answers for its CLIP path, both made with the real ViT-L/14 laion2B weights it downloads:
  * test/test_clip.py:6-12    <encode_text("ferrari f40"), embed(f40.jpg)> == 0.330654 (rtol/atol 1e-6)
  * test/clip_images/embeddings.pkl    the embeddings its pipeline stored for f40.jpg and micra.jpg
This script copies the two JPEG files (as byte arrays, decoded again by cv2 in the test exactly as the reference does) and
the stored vectors into tests/golden/clip_kat.npz.  The weights are not available offline, so the tests that use this file
(tests/test_kat_gated.py) run only when $CLEARCAM_B200_WEIGHTS holds CLIP-ViT-L-14-laion2B-s32B-b82K.safetensors.

    python oracle/make_golden_clip_kat.py
"""
import pickle
from pathlib import Path

import numpy as np

REF = Path("/root/reference/test/clip_images")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "clip_kat.npz"


def main():
    d = pickle.load(open(REF / "embeddings.pkl", "rb"))
    emb = {Path(k).name: np.asarray(v, np.float32).reshape(-1) for k, v in d["embeddings"].items()}
    np.savez_compressed(OUT, f40_jpg=np.frombuffer((REF / "f40.jpg").read_bytes(), np.uint8),
                        micra_jpg=np.frombuffer((REF / "micra.jpg").read_bytes(), np.uint8),
                        emb_f40=emb["f40.jpg"], emb_micra=emb["micra.jpg"], known_answer=np.float64(0.330654),
                        query=np.array("ferrari f40"))
    print("wrote", OUT, OUT.stat().st_size, "bytes; stored-embedding cosine f40 vs micra =", float(emb["f40.jpg"] @ emb["micra.jpg"]))


if __name__ == "__main__":
    main()

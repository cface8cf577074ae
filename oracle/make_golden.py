"""Generates the committed golden fixtures under tests/golden/ from the REFERENCE tree (run here, where
/root/reference exists; the GPU box only sees the fixtures).  Usage: python oracle/make_golden.py

  clip_tokens.json : token ids of fixed prompts from the reference's own utils/clip_tokenizer.py (imported, not copied)
                     -> pins clearcam_b200/utils/clip_tokenizer.py (tests/test_tokenizer_cpu.py)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PROMPTS = [
    "ferrari f40", "text here", "a photo of a cat", "person riding a bicycle at night", "Red TRUCK, parked!!",
    "it's 3 o'clock &amp; raining", "naïve café — déjà vu", "  multiple   spaces\tand\nnewlines ", "don't we've I'll he'd",
    "email@example.com 12345 #hashtag", "日本語のテキスト", "emoji 😀 test", "<start_of_text> weird <end_of_text>", "a" * 40,
    "white van with ladder on roof", "delivery driver carrying a box", "dog", "UPS", "x",
    "a man in a yellow jacket walking a large black dog past a parked silver car on a rainy evening",
]


def main():
    sys.path.insert(0, REF)
    from utils.clip_tokenizer import SimpleTokenizer  # the reference's tokenizer
    tok = SimpleTokenizer()
    out = {"source": "reference utils/clip_tokenizer.py SimpleTokenizer.encode", "prompts": PROMPTS,
           "ids": [tok.encode(p) for p in PROMPTS]}
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "clip_tokens.json"), "w") as f:
        json.dump(out, f, ensure_ascii=True, indent=0)
    print("wrote clip_tokens.json:", len(PROMPTS), "prompts")


if __name__ == "__main__":
    main()

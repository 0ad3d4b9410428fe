"""Fixtures for the file-format tests, generated in the build container from the reference's own data files
(they do not travel to the GPU box, so a small excerpt is committed):

  pretrained_2x4.json   wdf_py/diode_clipper/models/pretrained/'1N4148 (1U-1D)_2x4_pretrained_model.json' verbatim —
                        the Keras-written layout whose first entry is the InputLayer ({"type": "unknown", "weights": []})
  digilent_excerpt.csv  the 11 header lines of diode_dataset/1N4148/1up1down/10.0k_4.7nF.csv and its first 400 data rows

    python tests/golden/make_golden_io.py
"""
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    shutil.copyfile(os.path.join(REF, "wdf_py/diode_clipper/models/pretrained/1N4148 (1U-1D)_2x4_pretrained_model.json"), os.path.join(HERE, "pretrained_2x4.json"))
    with open(os.path.join(REF, "diode_dataset/1N4148/1up1down/10.0k_4.7nF.csv")) as f, open(os.path.join(HERE, "digilent_excerpt.csv"), "w") as g:
        for k, line in enumerate(f):
            if k >= 11 + 400:
                break
            g.write(line)
    print("wrote pretrained_2x4.json, digilent_excerpt.csv")

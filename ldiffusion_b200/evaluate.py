"""``python -m ldiffusion_b200.evaluate`` — the command line of the reference's ``evaluate.py:129-139``
(``python -m LDiffusion.evaluate``), same flags, scoring through the confusion-histogram kernel.

The package also exports the FUNCTION ``evaluate`` under this module's name (``from ldiffusion_b200 import
evaluate``).  Importing this submodule rebinds ``ldiffusion_b200.evaluate`` to the module, so the module is
made callable and forwards to the function: both spellings keep working in either import order.
"""
import argparse
import sys
import types

from .metrics import evaluate, frequency_weighted_iou, pixel_accuracy  # noqa: F401  (evaluate.py's public names)


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Evaluate segmentation results.")
    parser.add_argument("--image-dir", type=str, required=True, help="predicted images folder")
    parser.add_argument("--label-dir", type=str, required=True, help="labels folder")
    parser.add_argument("--num-classes", type=int, required=True, help="num-classes")
    parser.add_argument("--save-dir", type=str, default="./LDiffusion/eval/eval_report", help="results save folder")
    return parser.parse_args(argv)


class _CallableModule(types.ModuleType):
    def __call__(self, *args, **kwargs):
        return evaluate(*args, **kwargs)


sys.modules[__name__].__class__ = _CallableModule

if __name__ == "__main__":
    args = parse_args()
    evaluate(args.image_dir, args.label_dir, args.num_classes, args.save_dir)

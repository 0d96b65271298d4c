"""Mirror of the reference's data_handing/embeddings_related_generator.py (single input file).

Same four module-level functions, same CLI:
    python -m zsaac_b200.data_handing.embeddings_related_generator \
        --input_path data.pkl --output_path data_related.pkl --topnumber 5
"""
import argparse

from ..related_pipeline import load_data as _load_data
from ..related_pipeline import add_extension_flags, run_cli
from ..related_pipeline import process_data, save_data_to_hdf5  # noqa: F401  (re-exported)


def load_data(raw_path):
    """raw_path: str — one pickle holding a list of records (reference :9-17)."""
    return _load_data(raw_path)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--input_path', type=str, help="input path files")
    parser.add_argument('--output_path', type=str, help="output path files")
    parser.add_argument('--topnumber', type=int, default=5)
    add_extension_flags(parser)      # --gpus --exclude_self --[no-]rescore_fp32 --dtype --writer_procs --fast_pickle
    args = parser.parse_args(argv)
    run_cli(args, "zsaac_b200.data_handing.embeddings_related_generator", argv, load_data)


if __name__ == '__main__':
    main()

"""Mirror of the reference's data_handing/embeddings_related_generator_wavcaps.py (several input
files; `--input_path` takes one or more paths, reference :45).

    python -m zsaac_b200.data_handing.embeddings_related_generator_wavcaps \
        --input_path a.pkl b.pkl --output_path all_related.pkl --topnumber 5
"""
import argparse

from ..related_pipeline import load_data as _load_data
from ..related_pipeline import add_extension_flags, run_cli
from ..related_pipeline import process_data, save_data_to_hdf5  # noqa: F401  (re-exported)


def load_data(raw_path):
    """raw_path: list of str — pickles whose record lists are concatenated (reference :9-18)."""
    if isinstance(raw_path, (str, bytes)):
        # the reference iterates over raw_path, so a bare string would be read character by
        # character and fail in open(); fail with a clear message instead
        raise TypeError("embeddings_related_generator_wavcaps.load_data expects a list of paths")
    return _load_data(list(raw_path))


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--input_path', nargs='+', type=str)
    parser.add_argument('--output_path', type=str, help="output path files")
    parser.add_argument('--topnumber', type=int, default=5)
    add_extension_flags(parser)      # --gpus --exclude_self --[no-]rescore_fp32 --dtype --writer_procs --fast_pickle
    args = parser.parse_args(argv)
    run_cli(args, "zsaac_b200.data_handing.embeddings_related_generator_wavcaps", argv, load_data)


if __name__ == '__main__':
    main()

"""Mirror of the reference's data_handing/embeddings_related_generator_wavcaps.py (several input
files; `--input_path` takes one or more paths, reference :45).

    python -m zsaac_b200.data_handing.embeddings_related_generator_wavcaps \
        --input_path a.pkl b.pkl --output_path all_related.pkl --topnumber 5
"""
import argparse

from ..related_pipeline import load_data as _load_data
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
    # extension (not in the reference): pickle the output records in N forked processes
    parser.add_argument('--writer_procs', type=int, default=None)
    # extension: pickle tensors through numpy (same objects after pickle.load, ~3x faster)
    parser.add_argument('--fast_pickle', action='store_true', default=None)
    args = parser.parse_args(argv)
    valid_text_embs, all_data = load_data(args.input_path)
    processed_data_gen = process_data(valid_text_embs, all_data, args.topnumber)
    total_items = len(all_data)
    save_data_to_hdf5(processed_data_gen, args.output_path, total_items, workers=args.writer_procs,
                      fast_pickle=args.fast_pickle)


if __name__ == '__main__':
    main()

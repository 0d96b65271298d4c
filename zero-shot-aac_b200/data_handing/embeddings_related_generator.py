"""Mirror of the reference's data_handing/embeddings_related_generator.py (single input file).

Same four module-level functions, same CLI:
    python -m zsaac_b200.data_handing.embeddings_related_generator \
        --input_path data.pkl --output_path data_related.pkl --topnumber 5
"""
import argparse

from ..related_pipeline import load_data as _load_data
from ..related_pipeline import process_data, save_data_to_hdf5  # noqa: F401  (re-exported)


def load_data(raw_path):
    """raw_path: str — one pickle holding a list of records (reference :9-17)."""
    return _load_data(raw_path)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--input_path', type=str, help="input path files")
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

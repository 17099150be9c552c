#!/bin/bash
# Stage order and folder names of the reference's run.sh:60-70 (stage 0, deskew, is out of scope:
# point --input_folder at already oriented scans).
set -e
IN=${1:-0_oriented_images}
python 1_doclayout_bboxes.py --input_folder "$IN" --output_folder 1_doclayout_parsed
python 2_edge_box_filter.py --input_folder 1_doclayout_parsed --output_folder 2_edge_box_filtered
python 3_combine_grids.py --input_folder 2_edge_box_filtered --output_folder 3_combined_bboxes
python 4_extract_median_widths.py --input_folder 3_combined_bboxes/json --output_folder 4_medians_extracted
python 5_detect_column_centers.py --input_folder 3_combined_bboxes/json --median_folder 4_medians_extracted/json --output_folder 5_column_detection

"""Writes the SECOND golden picture the reference ships for its test run: the segmented grid AFTER the winning mutation.

Source: /root/reference/inst/extdata/output_png/Fasta_x_Fasta_y_all/Fasta_x_Fasta_y_all-7.png, drawn by
make_modified_grid_plot (R/plot_matrix_ggplot.R:261-297) from modify_gridmatrix (R/seqmodder.R:119-151), i.e. from
carry_out_compressed_sv (R/seqmodder.R:10-109) applied twice to the gridmatrix of ...-6.png: 4_12_inv, then 11_18_del
(README.md:135-139 mut_max).  The picture is an 11 x 12 grid with 13 filled cells, all blue (positive); its axes only
number the blocks, so what the picture holds per cell is the SIGN and the colour class of the legend
("Blocksize [kbp]": <= 100, <= 200, <= 300, <= 360).  Transcription: rows y1..y11 bottom to top, columns x1..x12.
Cell value written here = sign * upper bound of the colour class in bp.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
N_Y, N_X = 11, 12
# (y, x, colour class upper bound in kbp); every cell is blue = positive
cells = [(1, 1, 300), (2, 2, 100), (3, 3, 200), (4, 4, 200), (5, 5, 200), (6, 6, 100), (7, 7, 100), (7, 8, 100),
         (8, 8, 100), (8, 9, 100), (9, 10, 100), (10, 11, 100), (11, 12, 360)]
mat = [[0] * N_X for _ in range(N_Y)]
for y, x, c in cells:
    mat[y - 1][x - 1] = c * 1000
with open(os.path.join(HERE, "testrun_grid_after_mut.tsv"), "w") as f:
    for row in mat:
        f.write("\t".join(str(v) for v in row) + "\n")

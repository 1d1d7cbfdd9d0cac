"""Writes the test-run gridmatrix recovered from the reference's own golden picture.

Source: /root/reference/inst/extdata/output_png/Fasta_x_Fasta_y_all/Fasta_x_Fasta_y_all-6.png (axis labels
carry every block length, cell colours the signs) -- transcription recorded in SURVEY.md section 8c.
Cell value = sign * x-block length (R/bresenham.R:111-118); rows are y blocks, columns x blocks, exactly as
R/juliasolver.R:5-7 writes them.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
len_x = [290, 90, 110, 130, 30, 80, 10, 10, 10, 10, 110, 130, 90, 10, 10, 10, 80, 30, 360]
len_y = [290, 90, 110, 130, 110, 10, 10, 10, 80, 30, 360]
len_x = [v * 1000 for v in len_x]
len_y = [v * 1000 for v in len_y]
cells = [(1, 1, 1), (2, 2, 1), (2, 13, -1), (3, 3, 1), (4, 4, 1), (4, 12, -1), (5, 11, -1), (6, 10, -1),
         (6, 16, -1), (7, 8, -1), (7, 9, -1), (7, 15, -1), (8, 7, -1), (8, 8, -1), (8, 14, -1), (9, 6, -1),
         (10, 5, -1), (10, 18, 1), (11, 19, 1)]
mat = [[0] * len(len_x) for _ in len_y]
for y, x, s in cells:
    mat[y - 1][x - 1] = s * len_x[x - 1]
with open(os.path.join(HERE, "testrun_grid_mut.tsv"), "w") as f:
    for row in mat:
        f.write("\t".join(str(v) for v in row) + "\n")
with open(os.path.join(HERE, "testrun_grid_mut_lens.tsv"), "w") as f:
    f.write("\t".join(str(v) for v in len_y) + "\n")
    f.write("\t".join(str(v) for v in len_x) + "\n")

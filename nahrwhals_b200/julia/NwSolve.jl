# NwSolve.jl -- Julia bindings of libnwsolve.so (include/nwsolve.h, nwregion.h, nwgrid.h) for Julia-hosted pipelines.
#
# NOT EXECUTABLE IN THE BUILD IMAGE (no julia there).  The struct layouts mirror the C headers field by field; the
# Python bindings (nahrwhals_b200/_lib.py, region.py, grid.py) bind the same entry points and ARE tested.
#
#   using .NwSolve
#   ctx = NwSolve.Context()                                   # one per (process, GPU)
#   tsv = NwSolve.solve(ctx, aln, len_x, len_y; maxdepth = 3, maxdup = 2, maxwidth = 1000, minreport = 0.98, maxreport = 1000)
#   r   = NwSolve.analyse_region(ctx, mat, len_y, len_x, gridlines_x)     # r.mut_max, r.res_ref, r.res_max, r.table
#   g   = NwSolve.build_grid(ctx, tstart, tend, qstart, qend)             # g.gridlines_x, g.gridlines_y, g.mat
module NwSolve

const LIB = get(ENV, "NWSOLVE_LIB", joinpath(@__DIR__, "..", "csrc", "libnwsolve.so"))

struct NwOpts                      # struct nw_opts
    maxdepth::Int32
    maxdup::Int32
    maxwidth::Int64                # < 0 == nothing
    minreport::Float64
    maxreport::Int64               # < 0 == nothing
    keep_ids::Int32
    flags::Int32
end

struct NwRegionOpts                # struct nw_region_opts
    depth::Int32
    maxdup::Int32
    init_width::Float64
    minreport::Float64
    compression::Float64
    flags::Int32
    reserved::Int32
end

struct NwRegionSummary             # struct nw_region_summary
    res_ref::Float64
    res_max::Float64
    n_res_max::Int64
    mut_simulated::Int64
    mut_tested::Int64
    search_depth::Int32
    solver_called::Int32
    flipped::Int32
    flip_unsure::Int32
    cluttered::Int32
    n_rows::Int32
    n_mut::Int32
    width_reduced::Int32
    maxres::NTuple{15,Float64}
    prepass_candidates::Int64
    prepass_dp::Int64
    gpu_launches::Int32
    reserved::Int32
end

struct NwPaf                       # struct nw_paf
    tstart::Ptr{Float64}
    tend::Ptr{Float64}
    qstart::Ptr{Float64}
    qend::Ptr{Float64}
    n::Int64
end

last_error() = unsafe_string(ccall((:nw_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("nwsolve error $rc: $(last_error())")

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = -1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:nw_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(x -> (x.h != C_NULL && ccall((:nw_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), c)
        c
    end
end

"""aln: Int8 matrix n_x x n_y in {-1,0,1} (solver.jl:103-106 `transpose(sign.(raw))`).  Returns the report text
write_trajectories (solver.jl:747-753) would write."""
function solve(ctx::Context, aln::AbstractMatrix{Int8}, len_x::Vector{Int64}, len_y::Vector{Int64};
               maxdepth = 3, maxdup = 3, maxwidth = nothing, minreport = 0.95, maxreport = nothing, flags = 0)
    n_x, n_y = size(aln)
    a = collect(transpose(aln))                      # C side: x-major, aln[x * n_y + y]; Julia is column-major
    o = NwOpts(maxdepth, maxdup, something(maxwidth, -1), minreport, something(maxreport, -1), 0, flags)
    res = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:nw_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int8}, Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ref{NwOpts}, Ref{Ptr{Cvoid}}),
                ctx.h, a, n_x, n_y, len_x, len_y, Ref(o), res))
    data = Ref{Ptr{UInt8}}(C_NULL); len = Ref{Int64}(0)
    check(ccall((:nw_result_tsv, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{UInt8}}, Ref{Int64}), res[], data, len))
    tsv = unsafe_string(data[], len[])
    ccall((:nw_result_free, LIB), Cvoid, (Ptr{Cvoid},), res[])
    tsv
end

"""The solver stage of one region (R/wrapper_aln_and_analyse.R:176-191 + save_to_logfile's summary).
mat: Int64 matrix n_y x n_x of signed bp as R holds the gridmatrix."""
function analyse_region(ctx::Context, mat::AbstractMatrix{Int64}, len_y::Vector{Int64}, len_x::Vector{Int64},
                        gridlines_x::Vector{Float64}; depth = 3, maxdup = 2, init_width = 1000.0, minreport = 0.98,
                        compression = 1000.0, flags = 0)
    n_y, n_x = size(mat)
    g = collect(transpose(mat))                      # row-major by y for the C side
    o = NwRegionOpts(depth, maxdup, init_width, minreport, compression, flags, 0)
    tab = Ref{Ptr{Cvoid}}(C_NULL)
    summ = Ref{NwRegionSummary}()
    check(ccall((:nw_region_analyse, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int64}, Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32, Ref{NwRegionOpts},
                 Ref{Ptr{Cvoid}}, Ref{NwRegionSummary}),
                ctx.h, g, n_x, n_y, len_x, len_y, gridlines_x, length(gridlines_x), Ref(o), tab, summ))
    data = Ref{Ptr{UInt8}}(C_NULL); len = Ref{Int64}(0)
    check(ccall((:nw_table_tsv, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{UInt8}}, Ref{Int64}), tab[], data, len))
    table = unsafe_string(data[], len[])
    mut_max = unsafe_string(ccall((:nw_table_mut_max, LIB), Cstring, (Ptr{Cvoid},), tab[]))
    ccall((:nw_table_free, LIB), Cvoid, (Ptr{Cvoid},), tab[])
    s = summ[]
    (table = table, mut_max = mut_max, res_ref = s.res_ref, res_max = s.res_max, n_res_max = s.n_res_max,
     solver_called = s.solver_called != 0, summary = s)
end

"""Gridding of one region: alignments (paf columns) -> gridlines and gridmatrix (R/tsv_to_bitlocus.R:186-215)."""
function build_grid(ctx::Context, tstart::Vector{Float64}, tend::Vector{Float64}, qstart::Vector{Float64}, qend::Vector{Float64})
    GC.@preserve tstart tend qstart qend begin
        p = Ref(NwPaf(pointer(tstart), pointer(tend), pointer(qstart), pointer(qend), length(tstart)))
        out = Ref{Ptr{Cvoid}}(C_NULL); rc = Ref{Int32}(0)
        check(ccall((:nw_grid_build, LIB), Cint, (Ptr{Cvoid}, Ref{NwPaf}, Int64, Ref{Ptr{Cvoid}}, Ref{Int32}), ctx.h, p, 1, out, rc))
    end
    rc[] == 0 || error("gridding failed with status $(rc[])")   # -5: too many gridlines / alignments, -1: bad alignment
    nx = Ref{Int32}(0); ny = Ref{Int32}(0); cv = Ref{Int32}(0); gen = Ref{Int32}(0)
    check(ccall((:nw_grid_dims, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Int32}), out[], nx, ny, cv, gen))
    gx = Vector{Float64}(undef, nx[] + 1); gy = Vector{Float64}(undef, ny[] + 1)
    check(ccall((:nw_grid_lines, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), out[], gx, gy))
    m = Matrix{Float64}(undef, nx[], ny[])           # filled row-major by y == column-major (n_x, n_y); transposed below
    lx = Vector{Float64}(undef, nx[]); ly = Vector{Float64}(undef, ny[])
    check(ccall((:nw_grid_matrix, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), out[], m, lx, ly))
    ccall((:nw_grid_free, LIB), Cvoid, (Ptr{Cvoid},), out[])
    (gridlines_x = gx, gridlines_y = gy, mat = collect(transpose(m)), len_x = lx, len_y = ly, converged = cv[] != 0)
end

# ---- wrapper_paf_to_bitlocus (include/nwbitlocus.h): parsed PAF table -> accepted grid, with the reference's retry loop
struct NwPafCols                   # struct nw_paf_cols
    qlen::Ptr{Float64}; qstart::Ptr{Float64}; qend::Ptr{Float64}; strand::Ptr{Float64}; tlen::Ptr{Float64}
    tstart::Ptr{Float64}; tend::Ptr{Float64}; nmatch::Ptr{Float64}; alen::Ptr{Float64}; mapq::Ptr{Float64}
    n::Int64
end
struct NwBitlocusParams            # struct nw_bitlocus_params
    minlen::Float64; compression::Float64; chunklen::Float64
    max_n_alns::Int32; max_size_col_plus_rows::Int32; max_pairs::Int32; max_passes::Int32
end
struct NwBitlocusInfo              # struct nw_bitlocus_info
    minlen::Float64; compression::Float64
    passes::Int32; cluttered::Int32; n_alns::Int32; n_pairs::Int32
end

"paf: NamedTuple / Dict of Float64 vectors qlen, qstart, qend, strand (+1 / -1), tlen, tstart, tend, nmatch, alen, mapq.
Returns nothing for a cluttered PAF (as the reference does), else (gridlines_x, gridlines_y, mat, len_x, len_y, minlen,
compression, passes)."
function paf_to_bitlocus(ctx::Context, paf; minlen::Float64, compression::Float64, chunklen::Float64 = 10000.0,
                         max_n_alns::Integer = 150, max_size_col_plus_rows::Integer = 250)
    cols = [Vector{Float64}(paf[k]) for k in (:qlen, :qstart, :qend, :strand, :tlen, :tstart, :tend, :nmatch, :alen, :mapq)]
    g = Ref{Ptr{Cvoid}}(C_NULL); rc = Ref{Int32}(0)
    info = Ref(NwBitlocusInfo(0.0, 0.0, 0, 0, 0, -1))
    GC.@preserve cols begin
        c = Ref(NwPafCols(map(pointer, cols)..., length(cols[1])))
        prm = Ref(NwBitlocusParams(minlen, compression, chunklen, max_n_alns, max_size_col_plus_rows, 2000, 64))
        check(ccall((:nw_bitlocus_build, LIB), Cint,
                    (Ptr{Cvoid}, Ref{NwPafCols}, Int64, Ref{NwBitlocusParams}, Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Ref{NwBitlocusInfo}, Ref{Int32}),
                    ctx.h, c, 1, prm, g, C_NULL, info, rc))
    end
    rc[] == 0 || error("wrapper_paf_to_bitlocus failed with status $(rc[])")
    info[].cluttered != 0 && return nothing
    nx = Ref{Int32}(0); ny = Ref{Int32}(0); cv = Ref{Int32}(0); gen = Ref{Int32}(0)
    check(ccall((:nw_grid_dims, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Int32}), g[], nx, ny, cv, gen))
    gx = Vector{Float64}(undef, nx[] + 1); gy = Vector{Float64}(undef, ny[] + 1)
    check(ccall((:nw_grid_lines, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), g[], gx, gy))
    m = Matrix{Float64}(undef, nx[], ny[]); lx = Vector{Float64}(undef, nx[]); ly = Vector{Float64}(undef, ny[])
    check(ccall((:nw_grid_matrix, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), g[], m, lx, ly))
    ccall((:nw_grid_free, LIB), Cvoid, (Ptr{Cvoid},), g[])
    (gridlines_x = gx, gridlines_y = gy, mat = collect(transpose(m)), len_x = lx, len_y = ly,
     minlen = info[].minlen, compression = info[].compression, passes = Int(info[].passes))
end

# ---- regionfile mode from the alignments on, one call (include/nwpipeline.h)
struct NwLogMeta                   # struct nw_log_meta
    chr::Cstring; start::Cstring; stop::Cstring; samplename::Cstring
    y_seqname::Cstring; y_start::Cstring; y_end::Cstring; xpad::Cstring
    compression::Float64
    exceeds_x::Int32; exceeds_y::Int32; grid_inconsistency::Int32
end

"pafs: vector of PAF tables (as in paf_to_bitlocus); metas: vector of NamedTuples (chr, start, stop, samplename[, xpad]) of
strings.  Returns the rows of nahrwhals_res.tsv (nothing for a dropped region); logfile: also appended there."
function run_regions(ctx::Context, pafs::Vector, metas::Vector; minlen::Float64, compression::Float64,
                     chunklen::Float64 = 10000.0, depth::Integer = 3, maxdup::Integer = 2, init_width::Float64 = 1000.0,
                     minreport::Float64 = 0.98, logfile::Union{Nothing,String} = nothing)
    n = length(pafs)
    cols = [[Vector{Float64}(p[k]) for k in (:qlen, :qstart, :qend, :strand, :tlen, :tstart, :tend, :nmatch, :alen, :mapq)] for p in pafs]
    strs = [[String(m.chr), String(m.start), String(m.stop), String(m.samplename), String(get(m, :xpad, "1"))] for m in metas]
    text = Ref{Ptr{UInt8}}(C_NULL); len = Ref{Int64}(0)
    off = Vector{Int64}(undef, n + 1); st = Vector{Int32}(undef, max(n, 1))
    GC.@preserve cols strs begin
        pc = [NwPafCols(map(pointer, c)..., length(c[1])) for c in cols]
        pm = [NwLogMeta(pointer(s[1]), pointer(s[2]), pointer(s[3]), pointer(s[4]), C_NULL, C_NULL, C_NULL, pointer(s[5]),
                        NaN, 0, 0, 0) for s in strs]
        bp = Ref(NwBitlocusParams(minlen, compression, chunklen, 150, 250, 2000, 64))
        ro = Ref(NwRegionOpts(depth, maxdup, init_width, minreport, compression, 0, 0))
        check(ccall((:nw_regions_run, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{NwPafCols}, Ptr{NwLogMeta}, Int64, Ref{NwBitlocusParams}, Ref{NwRegionOpts}, Cstring,
                     Ref{Ptr{UInt8}}, Ref{Int64}, Ptr{Int64}, Ptr{Int32}),
                    ctx.h, pc, pm, n, bp, ro, logfile === nothing ? C_NULL : logfile, text, len, off, st))
    end
    blob = unsafe_string(text[], len[])
    ccall((:nw_free, LIB), Cvoid, (Ptr{Cvoid},), text[])
    [st[k] >= 0 ? chomp(blob[off[k]+1:off[k+1]]) : nothing for k in 1:n]
end

end # module

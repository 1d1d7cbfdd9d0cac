# solver.jl -- drop-in for NAHRwhals' inst/extdata/scripts/scripts_nw_main/solver.jl that forwards the search to
# libnwsolve.so (B200) through ccall.  Same command line as the reference script (its ArgParse table, lines 756-791):
#
#   julia solver.jl [-d D] [-u U] [-w W] [-O O] [-r r] [-R R] compressed_alignment compressed_lengths out_path
#
# NOT EXECUTABLE IN THE BUILD IMAGE (no julia there); it mirrors nahrwhals_b200/csrc/nwsolve_cli.cpp line by line,
# which *is* tested.  Point R at it with  params$julia_solver_path = "<repo>/nahrwhals_b200/julia/solver.jl".
# No package dependencies (the reference needs ArgParse / DelimitedFiles / ProgressMeter).

const LIBNWSOLVE = get(ENV, "NWSOLVE_LIB", joinpath(@__DIR__, "..", "csrc", "libnwsolve.so"))

# struct nw_opts of include/nwsolve.h (field order and C types must match)
struct NwOpts
    maxdepth::Int32
    maxdup::Int32
    maxwidth::Int64     # < 0 == nothing
    minreport::Float64
    maxreport::Int64    # < 0 == nothing
    keep_ids::Int32
    flags::Int32
end

function parse_cli(args::Vector{String})
    maxdepth, maxdup, maxwidth, minreport, maxreport = 3, 3, -1, 0.95, -1   # reference defaults
    positional = String[]
    i = 1
    while i <= length(args)
        a = args[i]
        if a in ("-d", "--maxdepth");      maxdepth = parse(Int, args[i+1]); i += 2
        elseif a in ("-u", "--maxdup");    maxdup = parse(Int, args[i+1]); i += 2
        elseif a in ("-w", "--maxwidth");  maxwidth = Int(parse(Float64, args[i+1])); i += 2
        elseif a in ("-O", "--maxoffset"); i += 2      # parsed and unused, as in the reference (slow_score ignores it)
        elseif a in ("-r", "--minreport"); minreport = parse(Float64, args[i+1]); i += 2
        elseif a in ("-R", "--maxreport"); maxreport = Int(parse(Float64, args[i+1])); i += 2
        else push!(positional, a); i += 1
        end
    end
    length(positional) == 3 || error("usage: solver.jl [options] compressed_alignment compressed_lengths out_path")
    NwOpts(maxdepth, maxdup, maxwidth, minreport, maxreport, 0, 0), positional
end

function main()::Cint
    opts, pos = parse_cli(ARGS)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:nw_ctx_create, LIBNWSOLVE), Cint, (Cint, Ref{Ptr{Cvoid}}), -1, ctx)
    if rc != 0
        println(stderr, "solver.jl: ", unsafe_string(ccall((:nw_last_error, LIBNWSOLVE), Cstring, ())))
        return 1
    end
    rc = ccall((:nw_solve_files, LIBNWSOLVE), Cint,
               (Ptr{Cvoid}, Cstring, Cstring, Cstring, Ref{NwOpts}, Ptr{Cvoid}),
               ctx[], pos[1], pos[2], pos[3], Ref(opts), C_NULL)
    if rc != 0
        println(stderr, "solver.jl: ", unsafe_string(ccall((:nw_last_error, LIBNWSOLVE), Cstring, ())))
    end
    ccall((:nw_ctx_destroy, LIBNWSOLVE), Cvoid, (Ptr{Cvoid},), ctx[])
    return rc == 0 ? 0 : 1
end

exit(main())

# julia/main_parking.jl -- the flow of AutonomousParking/main.jl:252-285 for Julia >= 1.6, on top of the shims of julia/OBCA.jl
# (ParkingDist / ParkingSignedDist / ParkingConstraints -> ccall -> libobca.so -> sm_100a kernels).
#
# The reference's main.jl is Julia 0.5/0.6 code (tic/toq, is_unix, linspace, ... -- SURVEY.md 8f-3) and so are its planner files.
# This runner keeps what main.jl does AFTER the planner: it reads the warm start that main.jl:215-248 would have produced
# (written as CSV by `python -m obca_b200.planner.export_warmstart DIR`, this repository's restatement of that part), calls the
# two NLP drivers with the reference's 17 positional arguments, runs the reference's acceptance test, prints the reference's
# summary and -- instead of the PyPlot animation of plotTraj.jl -- exports the trajectories as CSV.
#
#   julia julia/main_parking.jl DIR            (DIR as written by export_warmstart; results go to DIR/out_*.csv)
# Without DIR/xWS.csv (export_warmstart ... --no-plan) the warm start is planned here by libobca_planner.so (include/obca_planner.h).
#
# NOT executed in this repository's image (no Julia): written against include/obca.h and the tested ctypes twin obca_b200/_lib.py.
using DelimitedFiles
using Printf

include(joinpath(@__DIR__, "OBCA.jl"))

readmat(dir, name) = readdlm(joinpath(dir, name), ',', Float64)

function read_scalars(dir)
    d = Dict{String,Float64}()
    for ln in eachline(joinpath(dir, "scalars.csv"))
        k, v = split(ln, ',')
        d[String(k)] = parse(Float64, v)
    end
    return d
end

function export_solution(dir, tag, xp, up, timeScalep, lp, np_)
    writedlm(joinpath(dir, "out_$(tag)_x.csv"), permutedims(xp), ',')                 # (N+1) x 4: X, Y, psi, v
    writedlm(joinpath(dir, "out_$(tag)_u.csv"), permutedims(up), ',')                 # N x 2: delta, a
    writedlm(joinpath(dir, "out_$(tag)_timeScale.csv"), vec(timeScalep), ',')
    writedlm(joinpath(dir, "out_$(tag)_lambda.csv"), permutedims(lp), ',')
    writedlm(joinpath(dir, "out_$(tag)_mu.csv"), permutedims(np_), ',')
end

function main(dir)
    s = read_scalars(dir)
    N = Int(s["N"]); Ts = s["Ts"]; L = s["L"]; fixTime = Int(s["fixTime"]); nOb = Int(s["nOb"])
    x0 = readmat(dir, "x0.csv"); xF = readmat(dir, "xF.csv")                          # 1 x 4 row matrices, as in main.jl:108,213
    global ego = vec(readmat(dir, "ego.csv"))                                          # DualMultWS.jl:39 reads the global `ego`
    XYbounds = vec(readmat(dir, "XYbounds.csv"))
    vObMPC = Int.(readmat(dir, "vOb.csv"))                                             # half-space counts (main.jl:101)
    AOb = readmat(dir, "A.csv"); bOb = readmat(dir, "b.csv")                           # obstHrep (main.jl:252)
    if isfile(joinpath(dir, "xWS.csv"))
        path = readdlm(joinpath(dir, "path.csv"), ',', Float64; skipstart = 1)
        rx_sampled = path[:, 1]; ry_sampled = path[:, 2]; ryaw_sampled = path[:, 3]    # main.jl:237-239
        xWS = readmat(dir, "xWS.csv"); uWS = readmat(dir, "uWS.csv")                   # main.jl:247-248
    else
        # scenario files only (`export_warmstart DIR SCENARIO --no-plan`): plan here, with the native producer of libobca_planner.so
        # (Hybrid A* + veloSmooth + down-sampling = main.jl:215-248); the horizon N is the planner's (main.jl:244)
        w = plan_warm_start(vec(x0), vec(xF), Int(get(s, "scenario", 0.0)); Ts = Ts, L = L)
        w === nothing && error("Hybrid A*: cannot find a path")
        rx_sampled, ry_sampled, ryaw_sampled, xWS, uWS, N = w.rx, w.ry, w.ryaw, w.xWS, w.uWS, w.N
    end

    println("Parking using Distance Approach (A* warm start)")                         # main.jl:256-265
    xp20, up20, scaleTime20, exitflag20, time20, lp20, np20 =
        ParkingDist(x0, xF, N, Ts, L, ego, XYbounds, nOb, vObMPC, AOb, bOb, rx_sampled, ry_sampled, ryaw_sampled, fixTime, xWS, uWS)
    if exitflag20 == 1
        println("  --> Distance: SUCCESSFUL.")
        export_solution(dir, "dist", xp20, up20, scaleTime20, lp20, np20)
    else
        println("  --> WARNING: Problem could not be solved.")
    end
    ok20 = ParkingConstraints(x0, xF, N, Ts, L, ego, XYbounds, nOb, vObMPC, AOb, bOb, xp20, up20, lp20, np20, scaleTime20, fixTime, 0)

    println("Parking using Signed Distance Approach (A* warm start)")                  # main.jl:267-278
    xp10, up10, scaleTime10, exitflag10, time10, lp10, np10 =
        ParkingSignedDist(x0, xF, N, Ts, L, ego, XYbounds, nOb, vObMPC, AOb, bOb, rx_sampled, ry_sampled, ryaw_sampled, fixTime, xWS, uWS)
    if exitflag10 == 1
        println("  --> Signed Distance: SUCCESSFUL.")
        export_solution(dir, "signed", xp10, up10, scaleTime10, lp10, np10)
    else
        println("  --> WARNING: Problem could not be solved.")
    end
    ok10 = ParkingConstraints(x0, xF, N, Ts, L, ego, XYbounds, nOb, vObMPC, AOb, bOb, xp10, up10, lp10, np10, scaleTime10, fixTime, 1)

    println("********************* summary *********************")                     # main.jl:280-285
    @printf("  Time Distance approach: %.6f s   (ParkingConstraints: %d)\n", time20, ok20)
    @printf("  Time Signed Distance approach: %.6f s   (ParkingConstraints: %d)\n", time10, ok10)
    println("********************* DONE *********************")
    return exitflag20 == 1 && exitflag10 == 1
end

if abspath(PROGRAM_FILE) == @__FILE__
    main(length(ARGS) >= 1 ? ARGS[1] : ".") || exit(1)
end

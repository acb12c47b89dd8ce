# bench/ref_matvec.jl -- times the UNMODIFIED reference apply_H! (SpinDynamics.jl, Hamiltonian.jl:211-273) on the
# host cores, for machines that have Julia (the build container and the GPU boxes of this project do not, so
# bench.py's reference arm times the oracle port instead and says so: cpu_baseline.kind = "port").
#
#   julia -t auto --project=/path/to/SpinDynamics.jl bench/ref_matvec.jl [L=24] [reps=5]
#
# Prints one JSON line in the shape of bench.py's reference arm: same metric, same synthetic psi
# (psi[r] = 2 u(splitmix64(seed xor r)) - 1, seed 20261018, r = 0-based basis rank), threads stated.
using SpinDynamics, LinearAlgebra, Printf

function seeded_value(seed::UInt64, r::UInt64)
    z = (seed ⊻ r) + 0x9e3779b97f4a7c15
    z = (z ⊻ (z >> 30)) * 0xbf58476d1ce4e5b9
    z = (z ⊻ (z >> 27)) * 0x94d049bb133111eb
    z ⊻= z >> 31
    2.0 * (Float64(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0
end

function main()
    L = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 24
    reps = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 5
    model = XXZChain(L; Jxy=1.0, Jz=1.0, hz=0.0, nup=L ÷ 2, boundary=:open)
    N = length(model.states)
    ψ = [seeded_value(UInt64(20261018), UInt64(r - 1)) for r in 1:N]
    ψ ./= norm(ψ)
    out = similar(ψ)
    apply_H!(out, ψ, model)                                    # warm-up (compilation)
    t = @elapsed for _ in 1:reps
        apply_H!(out, ψ, model)
    end
    ms = 1e3 * t / reps
    @printf("{\"impl\": \"reference\", \"metric\": \"XXZ L=%d Sz=0 H.psi applies/s\", \"value\": %.6g, \"unit\": \"applies/s\", \"ms_per_step\": %.4f, \"dtype\": \"f64\", \"cpu_baseline\": {\"kind\": \"reference\", \"cores\": %d, \"sample\": \"SpinDynamics.apply_H! on XXZChain(%d; nup=%d), %d states, %d applies, Threads.nthreads()=%d\"}, \"checksum\": %.12e}\n",
            L, 1e3 / ms, ms, Threads.nthreads(), L, L ÷ 2, N, reps, Threads.nthreads(), dot(ψ, out))
end

main()

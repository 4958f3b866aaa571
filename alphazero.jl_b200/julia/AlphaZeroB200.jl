# AlphaZeroB200.jl -- thin Julia shim over libazb200.so (C ABI: include/azb200.h).
#
# NOT EXECUTED in the build image (no Julia toolchain there); it is deliberately mechanical: every function is one
# `ccall` plus array marshalling.  It adds a method at the seam `simulate_distributed(::Simulator, gspec, p)` used by
# `self_play_step!` (src/training.jl:284-286) so that `Scripts.train()` runs unchanged; games unknown to the library
# fall back to the reference's CPU path.
module AlphaZeroB200

using AlphaZero
using AlphaZero: GI, MCTS, Network, Trace, SimParams, MctsParams, SelfPlayParams, PLSchedule, ConstSchedule

const LIB = get(ENV, "AZB200_LIB", joinpath(@__DIR__, "..", "libazb200.so"))

struct CMctsParams            # az_mcts_params (src/params.jl:49-57)
  gamma::Cdouble; cpuct::Cdouble
  num_iters_per_turn::Int32; temperature_n::Int32
  dirichlet_noise_eps::Cdouble; dirichlet_noise_alpha::Cdouble; prior_temperature::Cdouble
  temperature_xs::NTuple{8,Int32}; temperature_ys::NTuple{8,Cdouble}
end
struct CSimParams             # az_sim_params (src/params.jl:92-101)
  num_games::Int32; num_workers::Int32; batch_size::Int32; fill_batches::Int32
  reset_every::Int32; alternate_colors::Int32; flip_probability::Cdouble
end
struct CResNetHP              # az_resnet_hp (src/networks/architectures/resnet.jl:30-37)
  num_blocks::Int32; num_filters::Int32; ksize::NTuple{2,Int32}
  num_policy_head_filters::Int32; num_value_head_filters::Int32; batch_norm_momentum::Cfloat
end

check(ctx, st) = st == 0 || error("azb200: " * unsafe_string(ccall((:az_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

function schedule_points(s)
  s isa ConstSchedule && return (Int32[0], Float64[s.value])
  s isa PLSchedule && return (Int32.(s.xs), Float64.(s.ys))
  error("unsupported temperature schedule")
end
function c_mcts_params(p::MctsParams)
  xs, ys = schedule_points(p.temperature)
  pad(v, z) = ntuple(i -> i <= length(v) ? v[i] : z, 8)
  CMctsParams(p.gamma, p.cpuct, p.num_iters_per_turn, length(xs), p.dirichlet_noise_ϵ, p.dirichlet_noise_α,
              p.prior_temperature, pad(xs, Int32(0)), pad(ys, 0.0))
end
c_sim_params(p::SimParams) = CSimParams(p.num_games, p.num_workers, p.batch_size, p.fill_batches,
  isnothing(p.reset_every) ? -1 : p.reset_every, p.alternate_colors, p.flip_probability)

# ---- state marshalling (formats: include/azb200.h header comment) -------------------------------------------
game_name(gspec) = occursin("connect", string(typeof(gspec))) ? "connect-four" :
                   occursin("tictactoe", lowercase(string(typeof(gspec)))) ? "tictactoe" :
                   occursin("mancala", lowercase(string(typeof(gspec)))) ? "mancala" : nothing
function state_from_bytes(::Val{Symbol("connect-four")}, b::AbstractVector{UInt8})
  (board = reshape(copy(b[1:42]), 7, 6) |> x -> typeof(GI.current_state(GI.init(Examples.games["connect-four"]))[:board])(x),
   curplayer = b[43])
end
state_to_bytes(::Val{Symbol("connect-four")}, s) = vcat(vec(Array(s.board)), s.curplayer)
# (tictactoe / mancala converters are the same two lines with their cell encodings)

# ---- network upload: Flux parameters in blob order (DESIGN.md "weight blob") --------------------------------
function flux_blob(nn)   # common, vhead, phead; Conv: W then b; BatchNorm: γ β μ σ²; Dense: W then b
  out = Float32[]
  for chain in (nn.common, nn.vhead, nn.phead), l in Flux.modules(chain)
    l isa Flux.Conv && (append!(out, vec(l.weight)); append!(out, l.bias))
    l isa Flux.BatchNorm && (append!(out, l.γ); append!(out, l.β); append!(out, l.μ); append!(out, l.σ²))
    l isa Flux.Dense && (append!(out, vec(l.weight)); append!(out, l.bias))
  end
  out
end

mutable struct Engine
  ctx::Ptr{Cvoid}; game::Int32; net::Ptr{Cvoid}
end
function Engine(gspec, nn; device = 0, ctx = nothing)   # pass `ctx = other.ctx` to put a second network on the same GPU context
  ctx = Ref{Ptr{Cvoid}}(ctx === nothing ? C_NULL : ctx)
  if ctx[] == C_NULL
    st = ccall((:az_ctx_create, LIB), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, ctx); check(C_NULL, st)
  end
  game = ccall((:az_game_lookup, LIB), Int32, (Cstring,), game_name(gspec))
  hp = Network.hyperparams(nn)
  chp = CResNetHP(hp.num_blocks, hp.num_filters, Int32.(hp.conv_kernel_size), hp.num_policy_head_filters,
                  hp.num_value_head_filters, hp.batch_norm_momentum)
  net = Ref{Ptr{Cvoid}}()
  check(ctx[], ccall((:az_net_create_resnet, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{CResNetHP}, Ptr{Ptr{Cvoid}}), ctx[], game, chp, net))
  blob = flux_blob(Network.to_cpu(nn))   # replaces Network.copy(bestnn; on_gpu=true, test_mode=true), src/training.jl:278
  check(ctx[], ccall((:az_net_load, LIB), Int32, (Ptr{Cvoid}, Ptr{Cfloat}, Int64), net[], blob, length(blob)))
  Engine(ctx[], game, net[])
end

# ---- the seam: simulate (src/simulations.jl:207-244) for the self-play simulator ---------------------------
function simulate_selfplay(e::Engine, gspec, params::SelfPlayParams; game_simulated, seed = rand(UInt64))
  mp, sp = c_mcts_params(params.mcts), c_sim_params(params.sim)
  h = Ref{Ptr{Cvoid}}()
  check(e.ctx, ccall((:az_selfplay_create, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{CMctsParams}, Ref{CSimParams}, UInt64, Ptr{Ptr{Cvoid}}), e.ctx, e.game, e.net, mp, sp, seed, h))
  check(e.ctx, ccall((:az_selfplay_start, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), h[], sp.num_games, 0))
  done, fin, seen = Ref{Int32}(0), Ref{Int32}(0), 0
  while fin[] == 0                              # az_selfplay_poll drives Handlers.game_played (no foreign-thread callback)
    check(e.ctx, ccall((:az_selfplay_poll, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), h[], done, fin))
    for _ in seen+1:done[]; game_simulated(); end
    seen = max(seen, done[]); sleep(0.01)
  end
  ns, ng = Ref{Int64}(0), Ref{Int64}(0)
  check(e.ctx, ccall((:az_selfplay_counts, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], ns, ng))
  A, SB = GI.num_actions(gspec), ccall((:az_game_state_bytes, LIB), Int32, (Int32,), e.game)
  states = Matrix{UInt8}(undef, SB, ns[]); pi = Matrix{Float32}(undef, A, ns[]); mask = Matrix{UInt8}(undef, A, ns[])
  z = Vector{Float32}(undef, ns[]); t = similar(z); gos = Vector{Int32}(undef, ns[]); rew = Vector{Float64}(undef, ns[])
  GC.@preserve states pi mask z t gos rew check(e.ctx, ccall((:az_selfplay_fetch, LIB), Int32,
        (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Float32}, Ptr{UInt8}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}),
        h[], states, pi, mask, z, t, gos, rew, C_NULL))
  edepth = Vector{Float64}(undef, ng[]); nodes = Vector{Int64}(undef, ng[]); moves = Vector{Int32}(undef, ng[])
  check(e.ctx, ccall((:az_selfplay_stats, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}),
        h[], edepth, nodes, moves, C_NULL))
  ccall((:az_selfplay_destroy, LIB), Int32, (Ptr{Cvoid},), h[])
  # rebuild the Vector of (trace, mem, edepth) that self_play_step! expects (src/training.jl:269-273,288-294)
  bytes_per_node = MCTS.memory_footprint_per_node(gspec)
  results, k = [], 1
  for g in 1:ng[]
    n = moves[g]
    sts = [state_from_bytes(Val(Symbol(game_name(gspec))), view(states, :, i)) for i in k:k+n-1]
    tr = Trace(sts[1])
    for i in 1:n
      next_state = i < n ? sts[i+1] : sts[i]    # the final state only matters for debug_trace
      push!(tr, Float64.(pi[mask[:, k+i-1] .== 1, k+i-1]), rew[k+i-1], next_state)   # compact π to the legal actions
    end
    push!(results, (trace = tr, mem = nodes[g] * bytes_per_node, edepth = edepth[g]))
    k += n
  end
  results
end

# pit_networks (src/training.jl:130-143) on the engine: a duel of two networks with the ArenaParams' mcts / sim settings.
# `contender` and `baseline` are Engines sharing one az_ctx (build the second with Engine(gspec, nn; ctx = first.ctx)).
function pit_networks(contender::Engine, baseline::Engine, gspec, params; game_simulated, seed = rand(UInt64))
  mp, sp = c_mcts_params(params.mcts), c_sim_params(params.sim)
  h = Ref{Ptr{Cvoid}}()
  check(contender.ctx, ccall((:az_selfplay_create_duel, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ref{CMctsParams}, Ref{CSimParams}, UInt64, Ptr{Ptr{Cvoid}}),
        contender.ctx, contender.game, contender.net, baseline.net, mp, sp, seed, h))
  check(contender.ctx, ccall((:az_selfplay_start, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), h[], sp.num_games, 0))
  done, fin, seen = Ref{Int32}(0), Ref{Int32}(0), 0
  while fin[] == 0
    check(contender.ctx, ccall((:az_selfplay_poll, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), h[], done, fin))
    for _ in seen+1:done[]; game_simulated(); end
    seen = max(seen, done[]); sleep(0.01)
  end
  rewards = Vector{Float64}(undef, sp.num_games); red = Ref{Float64}(0)
  GC.@preserve rewards check(contender.ctx, ccall((:az_selfplay_outcomes, LIB), Int32,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Int32}, Ptr{UInt8}, Ptr{Float64}), h[], params.mcts.gamma, rewards, C_NULL, C_NULL, red))
  ccall((:az_selfplay_destroy, LIB), Int32, (Ptr{Cvoid},), h[])
  return rewards, red[]                        # = rewards_and_redundancy(samples, gamma=params.mcts.gamma)
end

# Learning-side sample preparation on the GPU (src/learning.jl:38-51 after src/memory.jl:98-130): `samples` is the
# Vector{TrainingSample} of get_experience(env); returns the (W, X, A, P, V) Float32 tensors of convert_samples.
function prepare_samples(e::Engine, gspec, samples, wp::Int32; use_symmetries::Bool, merge::Bool = true)
  n, A = length(samples), GI.num_actions(gspec)
  SB = ccall((:az_game_state_bytes, LIB), Int32, (Int32,), e.game)
  states = Matrix{UInt8}(undef, SB, n); pi = zeros(Float64, A, n)
  for (i, s) in enumerate(samples)
    states[:, i] = state_to_bytes(Val(Symbol(game_name(gspec))), s.s)
    pi[GI.actions_mask(GI.init(gspec, s.s)), i] = s.π
  end
  z = Float64[s.z for s in samples]; t = Float64[s.t for s in samples]; cnt = Int32[s.n for s in samples]
  cur = Ref{Ptr{Cvoid}}()
  GC.@preserve states pi z t cnt check(e.ctx, ccall((:az_samples_from_host, LIB), Int32,
        (Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
        e.ctx, e.game, n, states, pi, z, t, cnt, cur))
  for (on, f) in ((use_symmetries, :az_samples_augment_with_symmetries), (merge, :az_samples_merge_by_state))
    on || continue
    nxt = Ref{Ptr{Cvoid}}()
    check(e.ctx, ccall((f, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), cur[], nxt))
    ccall((:az_samples_destroy, LIB), Int32, (Ptr{Cvoid},), cur[]); cur = nxt
  end
  m = Ref{Int64}(0); check(e.ctx, ccall((:az_samples_count, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), cur[], m))
  xdim = GI.state_dim(gspec)
  W = Matrix{Float32}(undef, 1, m[]); X = Array{Float32}(undef, xdim..., m[]); Am = Matrix{Float32}(undef, A, m[])
  P = similar(Am); V = Matrix{Float32}(undef, 1, m[])
  GC.@preserve W X Am P V check(e.ctx, ccall((:az_samples_convert, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), cur[], wp, W, X, Am, P, V))
  ccall((:az_samples_destroy, LIB), Int32, (Ptr{Cvoid},), cur[])
  return (; W, X, A = Am, P, V)
end

# ---- checkpoint interop (alphazero.jl_b200/checkpoint.py documents the format) --------------------------------------------
# Written next to the Serialization files of save_env (src/ui/session.jl:92-108) so the engine side can resume from a
# Julia session and vice versa.  JSON headers are hand-assembled to avoid a dependency beyond JSON3 (already a dependency).
function write_azb(path, gspec, nn)
  blob = flux_blob(Network.to_cpu(nn))
  header = JSON3.write((kind = "resnet", game = game_name(gspec), hyperparams = Network.hyperparams(nn),
                        num_params = length(blob), dtype = "float32", order = "flux"))
  open(path, "w") do io
    write(io, "AZB1"); write(io, htol(UInt32(sizeof(header)))); write(io, header); write(io, htol.(blob))
  end
end
function read_azb_blob(path)
  open(path, "r") do io
    String(read(io, 4)) == "AZB1" || error("not an AZB1 file: $path")
    n = ltoh(read(io, UInt32)); skip(io, n)
    return ltoh.(reinterpret(Float32, read(io)))
  end
end
# load a blob back into a Flux model (inverse of flux_blob): same traversal order
function import_weights!(nn, path)
  blob, q = read_azb_blob(path), 0
  take!(dst) = (copyto!(dst, reshape(view(blob, q+1:q+length(dst)), size(dst))); q += length(dst))
  for chain in (nn.common, nn.vhead, nn.phead), l in Flux.modules(chain)
    l isa Flux.Conv && (take!(l.weight); take!(l.bias))
    l isa Flux.BatchNorm && (take!(l.γ); take!(l.β); take!(l.μ); take!(l.σ²))
    l isa Flux.Dense && (take!(l.weight); take!(l.bias))
  end
  q == length(blob) || error("blob size mismatch")
  return nn
end
function write_azs(path, gspec, samples)   # get_experience(env) -> mem.azs
  n, A = length(samples), GI.num_actions(gspec)
  g = game_name(gspec)
  SB = ccall((:az_game_state_bytes, LIB), Int32, (Int32,), ccall((:az_game_lookup, LIB), Int32, (Cstring,), g))
  states = Matrix{UInt8}(undef, SB, n); pi = zeros(Float64, A, n)
  for (i, s) in enumerate(samples)
    states[:, i] = state_to_bytes(Val(Symbol(g)), s.s)
    pi[GI.actions_mask(GI.init(gspec, s.s)), i] = s.π
  end
  header = JSON3.write((game = g, num_samples = n, state_bytes = Int(SB), num_actions = A))
  open(path, "w") do io
    write(io, "AZS1"); write(io, htol(UInt32(sizeof(header)))); write(io, header)
    write(io, states); write(io, htol.(pi)); write(io, htol.(Float64[s.z for s in samples]))
    write(io, htol.(Float64[s.t for s in samples])); write(io, htol.(Int32[s.n for s in samples]))
  end
end
function export_session(env, dir)          # next to save_env(env, dir)
  write_azb(joinpath(dir, "bestnn.azb"), env.gspec, env.bestnn)
  write_azb(joinpath(dir, "curnn.azb"), env.gspec, env.curnn)
  write_azs(joinpath(dir, "mem.azs"), env.gspec, AlphaZero.get_experience(env))
end

end # module

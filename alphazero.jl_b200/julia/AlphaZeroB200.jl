# AlphaZeroB200.jl -- thin Julia shim over libazb200.so (C ABI: include/azb200.h).
#
# NOT EXECUTED in the build image (no Julia toolchain there): every function is one `ccall` plus array marshalling, and
# tests/test_julia_shim_cpu.py only checks what can be checked without Julia (every ccall names an exported symbol with
# the right number of arguments, blocks are balanced, every C struct mirrors include/azb200.h field for field).
#
# The seam (SURVEY.md 8b).  `self_play_step!` calls `simulate_distributed(simulator, gspec, params.sim; game_simulated)`
# (src/training.jl:284-286) and `pit_networks` / `evaluate_network` call `simulate(simulator, gspec, params.sim; ...)`
# (src/training.jl:137-139, 150-152).  Julia dispatches on the game-spec type, so loading this module ADDS the methods
#     simulate_distributed(::Simulator, ::Examples.ConnectFour.GameSpec, ::SimParams; game_simulated)   (and simulate)
# for the four games the library knows; no reference file changes and `Scripts.train("connect-four")` runs unchanged.
# Everything the engine cannot express (players other than MctsPlayer, NetworkPlayer (bare or under PlayerWithTemperature),
# TwoPlayers of those and TwoPlayers(such a player, MinMax.Player) -- Human, EpsilonGreedy, RandomPlayer --, oracles other than ResNet / SimpleNet / RolloutOracle / RandomOracle, a timeout instead of an iteration budget)
# falls back to the reference's own method through `invoke`.
module AlphaZeroB200

using AlphaZero
using AlphaZero: GI, MCTS, Network, NetLib, Examples, Trace, Simulator, SimParams, MctsParams, SelfPlayParams,
                 MctsPlayer, TwoPlayers, NetworkPlayer, PlayerWithTemperature, MinMax, AbstractGameSpec, AbstractSchedule,
                 PLSchedule, ConstSchedule
import Flux
import CUDA
import JSON3
import Distributed
using StaticArrays

const LIB = get(ENV, "AZB200_LIB", joinpath(@__DIR__, "..", "libazb200.so"))

# ---- C structs (field order and types = include/azb200.h) -----------------------------------------------------------
struct CMctsParams            # az_mcts_params (src/params.jl:49-57)
  gamma::Cdouble
  cpuct::Cdouble
  num_iters_per_turn::Int32
  temperature_n::Int32
  dirichlet_noise_eps::Cdouble
  dirichlet_noise_alpha::Cdouble
  prior_temperature::Cdouble
  temperature_xs::NTuple{8,Int32}
  temperature_ys::NTuple{8,Cdouble}
end
struct CMinMaxParams          # az_minmax_params (src/minmax.jl:72-81)
  depth::Int32
  amplify_rewards::Int32
  tau::Cdouble
  gamma::Cdouble
end
struct CSimParams             # az_sim_params (src/params.jl:92-101)
  num_games::Int32
  num_workers::Int32
  batch_size::Int32
  fill_batches::Int32
  reset_every::Int32
  alternate_colors::Int32
  flip_probability::Cdouble
end
struct CResNetHP              # az_resnet_hp (src/networks/architectures/resnet.jl:30-37)
  num_blocks::Int32
  num_filters::Int32
  conv_kernel_size::NTuple{2,Int32}
  num_policy_head_filters::Int32
  num_value_head_filters::Int32
  batch_norm_momentum::Cfloat
end
struct CSimpleNetHP           # az_simplenet_hp (src/networks/architectures/simplenet.jl:15-22)
  width::Int32
  depth_common::Int32
  depth_phead::Int32
  depth_vhead::Int32
  use_batch_norm::Int32
  batch_norm_momentum::Cfloat
end

last_error(ctx) = unsafe_string(ccall((:az_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
check(ctx, st) = st == 0 || error("azb200: " * last_error(ctx))

# ---- games: names = src/examples.jl:17-21, state byte formats = header comment of include/azb200.h ---------------------
const C4 = Examples.ConnectFour
const TTT = Examples.Tictactoe
const MAN = Examples.Mancala
const GW = Examples.GridWorld

game_name(::C4.GameSpec) = "connect-four"
game_name(::TTT.GameSpec) = "tictactoe"
game_name(::MAN.GameSpec) = "mancala"
game_name(gspec::AbstractGameSpec) = gspec === Examples.games["grid-world"] || typeof(gspec) == typeof(Examples.games["grid-world"]) ? "grid-world" : nothing

# connect-four (43 B): cells[col + 7*row] (the memory order of SMatrix{7,6,UInt8}), curplayer
state_to_bytes(::C4.GameSpec, s) = vcat(vec(Array(s.board)), UInt8(s.curplayer))
state_from_bytes(::C4.GameSpec, b) = (board = C4.Board(b[1:42]), curplayer = UInt8(b[43]))
# tictactoe (10 B): cells[pos] in {0 empty, 1 white, 2 black} (Cell = Union{Nothing, Bool}, WHITE = true), curplayer {1, 2}
ttt_cell_to_byte(c) = isnothing(c) ? 0x00 : (c ? 0x01 : 0x02)
ttt_byte_to_cell(x) = x == 0x00 ? nothing : (x == 0x01)
state_to_bytes(::TTT.GameSpec, s) = vcat(UInt8[ttt_cell_to_byte(c) for c in s.board], s.curplayer ? 0x01 : 0x02)
state_from_bytes(::TTT.GameSpec, b) = (board = TTT.Board(TTT.Cell[ttt_byte_to_cell(x) for x in b[1:9]]), curplayer = (b[10] == 0x01))
# mancala (15 B): stores[2], houses[(player-1) + 2*(num-1)] (the memory order of SMatrix{2,6,UInt8}), curplayer {1, 2}
state_to_bytes(::MAN.GameSpec, s) = vcat(UInt8.(collect(s.board.stores)), vec(Array(s.board.houses)), UInt8(s.curplayer))
state_from_bytes(::MAN.GameSpec, b) =
  (board = MAN.Board(SVector{2,UInt8}(b[1], b[2]), SMatrix{2,6,UInt8,12}(b[3:14])), curplayer = Int(b[15]))
# grid-world (2 B): x, y in 1..10 (the state of the CommonRLInterface wrapper is the SVector{2,Int} observation)
state_to_bytes(gspec::AbstractGameSpec, s) = UInt8[s[1], s[2]]
state_from_bytes(gspec::AbstractGameSpec, b) = SVector{2,Int}(Int(b[1]), Int(b[2]))

# ---- parameters -----------------------------------------------------------------------------------------------------------
function schedule_points(s::AbstractSchedule)
  s isa ConstSchedule && return (Int32[0], Float64[s.value])
  s isa PLSchedule && return (Int32.(s.xs), Float64.(s.ys))
  return nothing            # StepSchedule / CyclicSchedule: not expressible as <= 8 linear pieces in general
end
function c_mcts_params(gamma, cpuct, niters, eps, alpha, prior_temperature, schedule)
  pts = schedule_points(schedule)
  (isnothing(pts) || length(pts[1]) > 8) && return nothing
  xs, ys = pts
  pad(v, z) = ntuple(i -> i <= length(v) ? v[i] : z, 8)
  return CMctsParams(gamma, cpuct, niters, length(xs), eps, alpha, prior_temperature, pad(xs, Int32(0)), pad(ys, 0.0))
end
c_mcts_params(p::MctsParams) = c_mcts_params(p.gamma, p.cpuct, p.num_iters_per_turn, p.dirichlet_noise_ϵ, p.dirichlet_noise_α,
                                             p.prior_temperature, p.temperature)
# the parameters a constructed MctsPlayer carries (src/play.jl:156-181): lets the seam read them back from the player the
# Simulator's `make_player` closure builds, whatever MctsParams object the caller captured in that closure
function c_mcts_params(pl::MctsPlayer)
  isnothing(pl.timeout) || return nothing
  e = pl.mcts
  return c_mcts_params(e.gamma, e.cpuct, pl.niters, e.noise_ϵ, e.noise_α, e.prior_temperature, pl.τ)
end
# Benchmark.NetworkOnly (src/benchmark.jl:161-176) = PlayerWithTemperature(NetworkPlayer(nn), ConstSchedule(τ)); a bare
# NetworkPlayer plays at temperature 1 (src/play.jl:37-39).  num_iters_per_turn = 0 selects the network-only player in the
# library (include/azb200.h); the search fields are inert.
c_mcts_params(pl::NetworkPlayer) = c_mcts_params(1.0, 1.0, 0, 0.0, 1.0, 1.0, ConstSchedule(1.0))
function c_mcts_params(pl::PlayerWithTemperature)
  pl.player isa NetworkPlayer || return nothing
  return c_mcts_params(1.0, 1.0, 0, 0.0, 1.0, 1.0, pl.temperature)
end
c_mcts_params(::Any) = nothing
engine_player(pl) = pl isa MctsPlayer || pl isa NetworkPlayer || (pl isa PlayerWithTemperature && pl.player isa NetworkPlayer)
c_sim_params(p::SimParams, num_games) = CSimParams(num_games, p.num_workers, p.batch_size, p.fill_batches ? 1 : 0,
  isnothing(p.reset_every) ? -1 : p.reset_every, p.alternate_colors ? 1 : 0, p.flip_probability)

# ---- network upload: Flux parameters in blob order (include/azb200.h "az_net_num_params") -------------------------------
function flux_blob(nn)   # common, vhead, phead; Conv: W then b; BatchNorm: γ β μ σ²; Dense: W then b
  out = Float32[]
  for chain in (nn.common, nn.vhead, nn.phead)
    for l in Flux.modules(chain)
      if l isa Flux.Conv
        append!(out, vec(l.weight)); append!(out, l.bias)
      elseif l isa Flux.BatchNorm
        append!(out, l.γ); append!(out, l.β); append!(out, l.μ); append!(out, l.σ²)
      elseif l isa Flux.Dense
        append!(out, vec(l.weight)); append!(out, l.bias)
      end
    end
  end
  return out
end

# oracles the engine has a device implementation of: the two NetLib networks, MCTS.RolloutOracle (src/mcts.jl:27-60, the oracle
# of Benchmark.MctsRollouts) and MCTS.RandomOracle (src/mcts.jl:62-72)
# the same blob built ON THE GPU for a network that already lives there (Network.on_gpu): one CuArray, no host copy
function flux_blob_device(nn)
  parts = CUDA.CuArray{Float32}[]
  for chain in (nn.common, nn.vhead, nn.phead)
    for l in Flux.modules(chain)
      if l isa Flux.Conv
        push!(parts, vec(l.weight)); push!(parts, vec(l.bias))
      elseif l isa Flux.BatchNorm
        push!(parts, vec(l.γ)); push!(parts, vec(l.β)); push!(parts, vec(l.μ)); push!(parts, vec(l.σ²))
      elseif l isa Flux.Dense
        push!(parts, vec(l.weight)); push!(parts, vec(l.bias))
      end
    end
  end
  return reduce(vcat, parts)
end

supported_network(nn) = nn isa NetLib.ResNet || nn isa NetLib.SimpleNet || nn isa MCTS.RolloutOracle || nn isa MCTS.RandomOracle

mutable struct Engine
  ctx::Ptr{Cvoid}
  game::Int32
  net::Ptr{Cvoid}
  owns_ctx::Bool
end
const CONTEXTS = Dict{Int,Ptr{Cvoid}}()     # one az_ctx per GPU per process
function context(device::Int)
  get!(CONTEXTS, device) do
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:az_ctx_create, LIB), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, ctx)
    st == 0 || error("azb200: " * last_error(C_NULL))
    ctx[]
  end
end
# the device of this process: worker w of `Distributed.workers()` drives GPU (index of w) - 1, the master GPU 0
default_device() = max(0, something(findfirst(==(Distributed.myid()), Distributed.workers()), 1) - 1)

function Engine(gspec, nn; device = default_device())
  ctx = context(device)
  game = ccall((:az_game_lookup, LIB), Int32, (Cstring,), game_name(gspec))
  game >= 0 || error("azb200: unknown game")
  net = Ref{Ptr{Cvoid}}(C_NULL)
  if nn isa MCTS.RolloutOracle     # playout draws come from the engine's Philox stream keyed by (seed, state, ply)
    check(ctx, ccall((:az_net_create_rollout, LIB), Int32, (Ptr{Cvoid}, Int32, Cdouble, UInt64, Ptr{Ptr{Cvoid}}), ctx, game, nn.gamma, rand(UInt64), net))
    return Engine(ctx, game, net[], false)
  elseif nn isa MCTS.RandomOracle
    check(ctx, ccall((:az_net_create_oracle, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Ptr{Cvoid}}), ctx, Int32(0), game, net))
    return Engine(ctx, game, net[], false)
  end
  hp = Network.hyperparams(nn)
  if nn isa NetLib.ResNet
    chp = CResNetHP(hp.num_blocks, hp.num_filters, (Int32(hp.conv_kernel_size[1]), Int32(hp.conv_kernel_size[2])),
                    hp.num_policy_head_filters, hp.num_value_head_filters, hp.batch_norm_momentum)
    check(ctx, ccall((:az_net_create_resnet, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{CResNetHP}, Ptr{Ptr{Cvoid}}), ctx, game, chp, net))
  else
    chp = CSimpleNetHP(hp.width, hp.depth_common, hp.depth_phead, hp.depth_vhead, hp.use_batch_norm ? 1 : 0, hp.batch_norm_momentum)
    check(ctx, ccall((:az_net_create_simplenet, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{CSimpleNetHP}, Ptr{Ptr{Cvoid}}), ctx, game, chp, net))
  end
  # replaces Network.copy(bestnn; on_gpu=true, test_mode=true), src/training.jl:278
  if Network.on_gpu(nn) && CUDA.deviceid(CUDA.device()) == device
    dblob = flux_blob_device(nn)          # parameters never leave the GPU: the library folds BatchNorm on the device
    CUDA.synchronize()
    GC.@preserve dblob check(ctx, ccall((:az_net_load_device, LIB), Int32, (Ptr{Cvoid}, CUDA.CuPtr{Cfloat}, Int64), net[], dblob, length(dblob)))
  else
    blob = flux_blob(Network.to_cpu(nn))
    GC.@preserve blob check(ctx, ccall((:az_net_load, LIB), Int32, (Ptr{Cvoid}, Ptr{Cfloat}, Int64), net[], blob, length(blob)))
  end
  return Engine(ctx, game, net[], false)
end
close!(e::Engine) = (ccall((:az_net_destroy, LIB), Int32, (Ptr{Cvoid},), e.net); nothing)

# ---- running one batch of games on this process's GPU -----------------------------------------------------------------
function poll_until_finished(ctx, h, game_simulated)
  done, fin, seen = Ref{Int32}(0), Ref{Int32}(0), 0
  while fin[] == 0                              # az_selfplay_poll drives game_simulated (no foreign-thread callback)
    check(ctx, ccall((:az_selfplay_poll, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), h, done, fin))
    for _ in (seen + 1):done[]
      game_simulated()
    end
    seen = max(seen, Int(done[]))
    fin[] == 0 && sleep(0.005)
  end
  for _ in (seen + 1):done[]
    game_simulated()
  end
end

# rebuild the reference's Trace objects (src/trace.jl:17-47): n + 1 states, n compact policies, n rewards per game
function fetch_traces(ctx, h, gspec)
  ns, ng = Ref{Int64}(0), Ref{Int64}(0)
  check(ctx, ccall((:az_selfplay_counts, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h, ns, ng))
  game = ccall((:az_game_lookup, LIB), Int32, (Cstring,), game_name(gspec))
  A, SB = GI.num_actions(gspec), Int(ccall((:az_game_state_bytes, LIB), Int32, (Int32,), game))
  states = Matrix{UInt8}(undef, SB, ns[]); pi = Matrix{Float32}(undef, A, ns[]); mask = Matrix{UInt8}(undef, A, ns[])
  z = Vector{Float32}(undef, ns[]); t = Vector{Float32}(undef, ns[]); gos = Vector{Int32}(undef, ns[]); rew = Vector{Float64}(undef, ns[])
  GC.@preserve states pi mask z t gos rew check(ctx, ccall((:az_selfplay_fetch, LIB), Int32,
        (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Float32}, Ptr{UInt8}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}),
        h, states, pi, mask, z, t, gos, rew, C_NULL))
  edepth = Vector{Float64}(undef, ng[]); nodes = Vector{Int64}(undef, ng[]); moves = Vector{Int32}(undef, ng[])
  GC.@preserve edepth nodes moves check(ctx, ccall((:az_selfplay_stats, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}), h, edepth, nodes, moves, C_NULL))
  finals = Matrix{UInt8}(undef, SB, ng[]); flipped = Vector{Int32}(undef, ng[])
  GC.@preserve finals flipped check(ctx, ccall((:az_selfplay_outcomes, LIB), Int32,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Int32}, Ptr{UInt8}, Ptr{Float64}), h, 1.0, C_NULL, flipped, finals, C_NULL))
  traces = Vector{Any}(undef, ng[])
  k = 1
  for g in 1:ng[]
    n = Int(moves[g])
    sts = [state_from_bytes(gspec, states[:, i]) for i in k:(k + n - 1)]
    push!(sts, state_from_bytes(gspec, finals[:, g]))
    tr = Trace(sts[1])
    for i in 1:n
      col = k + i - 1
      push!(tr, Float64.(pi[mask[:, col] .== 0x01, col]), rew[col], sts[i + 1])   # policy compact over the legal actions
    end
    traces[g] = tr
    k += n
  end
  return traces, edepth, nodes, flipped
end

function run_selfplay(gspec, nn, mp::CMctsParams, sp::CSimParams, first_game::Int; game_simulated, seed)
  e = Engine(gspec, nn)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(e.ctx, ccall((:az_selfplay_create, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{CMctsParams}, Ref{CSimParams}, UInt64, Ptr{Ptr{Cvoid}}), e.ctx, e.game, e.net, mp, sp, seed, h))
  check(e.ctx, ccall((:az_selfplay_start, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), h[], sp.num_games, first_game))
  poll_until_finished(e.ctx, h[], game_simulated)
  out = fetch_traces(e.ctx, h[], gspec)
  ccall((:az_selfplay_destroy, LIB), Int32, (Ptr{Cvoid},), h[])
  close!(e)
  return out
end
function run_duel(gspec, nn_white, nn_black, mp::CMctsParams, mp_black::CMctsParams, sp::CSimParams, first_game::Int; game_simulated, seed)
  w = Engine(gspec, nn_white)
  b = Engine(gspec, nn_black)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(w.ctx, ccall((:az_selfplay_create_duel_players, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{CMctsParams}, Ptr{Cvoid}, Ref{CMctsParams}, Ref{CSimParams}, UInt64, Ptr{Ptr{Cvoid}}),
        w.ctx, w.game, w.net, mp, b.net, mp_black, sp, seed, h))
  check(w.ctx, ccall((:az_selfplay_start, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), h[], sp.num_games, first_game))
  poll_until_finished(w.ctx, h[], game_simulated)
  out = fetch_traces(w.ctx, h[], gspec)
  ccall((:az_selfplay_destroy, LIB), Int32, (Ptr{Cvoid},), h[])
  close!(w); close!(b)
  return out
end

function run_duel_minmax(gspec, nn_white, mp::CMctsParams, mm::CMinMaxParams, sp::CSimParams, first_game::Int; game_simulated, seed)
  w = Engine(gspec, nn_white)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(w.ctx, ccall((:az_selfplay_create_duel_minmax, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{CMctsParams}, Ref{CMinMaxParams}, Ref{CSimParams}, UInt64, Ptr{Ptr{Cvoid}}),
        w.ctx, w.game, w.net, mp, mm, sp, seed, h))
  check(w.ctx, ccall((:az_selfplay_start, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), h[], sp.num_games, first_game))
  poll_until_finished(w.ctx, h[], game_simulated)
  out = fetch_traces(w.ctx, h[], gspec)
  ccall((:az_selfplay_destroy, LIB), Int32, (Ptr{Cvoid},), h[])
  close!(w)
  return out
end

# ---- the seam ------------------------------------------------------------------------------------------------------------
# What the engine can run: one MctsPlayer or NetworkPlayer, or TwoPlayers of two such players (each with its own parameters
# and oracle: ResNet, SimpleNet, MCTS.RolloutOracle, MCTS.RandomOracle), measured by self_play_measurements (src/training.jl:269-273) or
# record_trace (src/simulations.jl:195).
# the oracle a constructed player consults: pit_networks hands make_player a tuple of two networks (src/training.jl:131-139),
# Benchmark.run ONE network from which both players are instantiated (src/benchmark.jl:82-93: Benchmark.MctsRollouts ignores
# it and brings a RolloutOracle) -- reading it off the player covers both
player_oracle(pl::MctsPlayer) = pl.mcts.oracle
player_oracle(pl::NetworkPlayer) = pl.network
player_oracle(pl::PlayerWithTemperature) = player_oracle(pl.player)
player_oracle(::Any) = nothing
function plan(simulator::Simulator, gspec)
  isnothing(game_name(gspec)) && return nothing
  oracles = simulator.make_oracles()
  player = simulator.make_player(oracles)
  bpn = MCTS.memory_footprint_per_node(gspec)
  if engine_player(player) && supported_network(player_oracle(player))
    mp = c_mcts_params(player)
    isnothing(mp) && return nothing
    return (kind = :single, nets = (player_oracle(player),), mp = mp, mp_black = mp, mm = nothing, bytes_per_node = bpn)
  elseif player isa TwoPlayers && engine_player(player.white) && supported_network(player_oracle(player.white))
    mpw = c_mcts_params(player.white)
    isnothing(mpw) && return nothing
    b = player.black
    if b isa MinMax.Player   # Benchmark.MinMaxTS (src/benchmark.jl:178-196): searched on the device, no oracle
      (1 <= b.depth <= 8 && !isnothing(game_heuristic(gspec))) || return nothing
      mm = CMinMaxParams(b.depth, b.amplify_rewards ? 1 : 0, b.τ, b.gamma)
      return (kind = :minmax, nets = (player_oracle(player.white),), mp = mpw, mp_black = mpw, mm = mm, bytes_per_node = bpn)
    elseif engine_player(b) && supported_network(player_oracle(b))
      mpb = c_mcts_params(b)
      isnothing(mpb) && return nothing
      return (kind = :duel, nets = (player_oracle(player.white), player_oracle(b)), mp = mpw, mp_black = mpb, mm = nothing, bytes_per_node = bpn)
    end
  end
  return nothing
end
game_heuristic(gspec) = game_name(gspec) == "grid-world" ? nothing : true   # GI.heuristic_value of grid-world is the constant 0

function measure(simulator::Simulator, pl, trace, colors_flipped, edepth, nodes)
  if simulator.measure === AlphaZero.self_play_measurements
    return (trace = trace, mem = nodes * pl.bytes_per_node, edepth = edepth)
  elseif simulator.measure === AlphaZero.record_trace
    return (trace = trace, colors_flipped = colors_flipped)
  end
  return nothing
end
measurable(simulator::Simulator) = simulator.measure === AlphaZero.self_play_measurements || simulator.measure === AlphaZero.record_trace

# `simulate` on this process's GPU: games first_game .. first_game + num_games - 1 (global indices key the RNG streams)
function simulate_on_gpu(simulator::Simulator, gspec, p::SimParams, pl, num_games::Int, first_game::Int; game_simulated,
                         seed = rand(UInt64))
  sp = c_sim_params(p, num_games)
  traces, edepth, nodes, flipped =
    pl.kind == :single ?
      run_selfplay(gspec, pl.nets[1], pl.mp, sp, first_game; game_simulated = game_simulated, seed = seed) :
    pl.kind == :minmax ?
      run_duel_minmax(gspec, pl.nets[1], pl.mp, pl.mm, sp, first_game; game_simulated = game_simulated, seed = seed) :
      run_duel(gspec, pl.nets[1], pl.nets[2], pl.mp, pl.mp_black, sp, first_game; game_simulated = game_simulated, seed = seed)
  return [measure(simulator, pl, traces[g], flipped[g] != 0, edepth[g], nodes[g]) for g in 1:length(traces)]
end

for S in (:(C4.GameSpec), :(TTT.GameSpec), :(MAN.GameSpec), :(typeof(Examples.games["grid-world"])))
  @eval begin
    # simulate (src/simulations.jl:207-244): pit_networks / evaluate_network / Benchmark.run call this one directly
    function AlphaZero.simulate(simulator::Simulator, gspec::$S, p::SimParams; game_simulated)
      pl = measurable(simulator) ? plan(simulator, gspec) : nothing
      if isnothing(pl)
        return invoke(AlphaZero.simulate, Tuple{Simulator, AbstractGameSpec, SimParams}, simulator, gspec, p; game_simulated = game_simulated)
      end
      return simulate_on_gpu(simulator, gspec, p, pl, p.num_games, 0; game_simulated = game_simulated)
    end
    # simulate_distributed (src/simulations.jl:252-290): one Distributed worker per GPU, num_each / rem split (:268, :277),
    # progress through the same RemoteChannel pattern, results concatenated in worker order (reduce(vcat, results), :289)
    function AlphaZero.simulate_distributed(simulator::Simulator, gspec::$S, p::SimParams; game_simulated)
      pl = measurable(simulator) ? plan(simulator, gspec) : nothing
      if isnothing(pl)
        return invoke(AlphaZero.simulate_distributed, Tuple{Simulator, AbstractGameSpec, SimParams}, simulator, gspec, p;
                      game_simulated = game_simulated)
      end
      workers = Distributed.workers()
      if length(workers) == 1 && workers[1] == Distributed.myid()
        return simulate_on_gpu(simulator, gspec, p, pl, p.num_games, 0; game_simulated = game_simulated)
      end
      chan = Distributed.RemoteChannel(() -> Channel{Nothing}(p.num_games))
      counter = @async for _ in 1:p.num_games
        take!(chan)
        game_simulated()
      end
      remote_game_simulated() = put!(chan, nothing)
      num_each, rem = divrem(p.num_games, length(workers))
      @assert num_each >= 1
      seed = rand(UInt64)
      tasks = map(enumerate(workers)) do (i, w)
        count = i == 1 ? num_each + rem : num_each
        first = i == 1 ? 0 : num_each * (i - 1) + rem
        Distributed.@spawnat w AlphaZeroB200.simulate_on_gpu(simulator, gspec, p, AlphaZeroB200.plan(simulator, gspec), count, first;
                                                            game_simulated = remote_game_simulated, seed = seed)
      end
      results = fetch.(tasks)
      wait(counter)
      return reduce(vcat, results)
    end
  end
end

# ---- learning-side sample preparation on the GPU (src/learning.jl:38-51 after src/memory.jl:98-130) ---------------------
# `samples` is the Vector{TrainingSample} of get_experience(env); returns the (W, X, A, P, V) Float32 tensors of convert_samples.
function prepare_samples(gspec, samples, wp::Integer; use_symmetries::Bool, merge::Bool = true, device = default_device())
  ctx = context(device)
  game = ccall((:az_game_lookup, LIB), Int32, (Cstring,), game_name(gspec))
  n, A = length(samples), GI.num_actions(gspec)
  SB = Int(ccall((:az_game_state_bytes, LIB), Int32, (Int32,), game))
  states = Matrix{UInt8}(undef, SB, n); pi = zeros(Float64, A, n)
  for (i, s) in enumerate(samples)
    states[:, i] = state_to_bytes(gspec, s.s)
    pi[GI.actions_mask(GI.init(gspec, s.s)), i] = s.π
  end
  z = Float64[s.z for s in samples]; t = Float64[s.t for s in samples]; cnt = Int32[s.n for s in samples]
  cur = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve states pi z t cnt check(ctx, ccall((:az_samples_from_host, LIB), Int32,
        (Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Ptr{Cvoid}}),
        ctx, game, n, states, pi, z, t, cnt, cur))
  if use_symmetries
    nxt = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:az_samples_augment_with_symmetries, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), cur[], nxt))
    ccall((:az_samples_destroy, LIB), Int32, (Ptr{Cvoid},), cur[])
    cur = nxt
  end
  if merge
    nxt = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:az_samples_merge_by_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), cur[], nxt))
    ccall((:az_samples_destroy, LIB), Int32, (Ptr{Cvoid},), cur[])
    cur = nxt
  end
  m = Ref{Int64}(0)
  check(ctx, ccall((:az_samples_count, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), cur[], m))
  xdim = GI.state_dim(gspec)
  W = Matrix{Float32}(undef, 1, m[]); X = Array{Float32}(undef, xdim..., m[]); Am = Matrix{Float32}(undef, A, m[])
  P = Matrix{Float32}(undef, A, m[]); V = Matrix{Float32}(undef, 1, m[])
  GC.@preserve W X Am P V check(ctx, ccall((:az_samples_convert, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), cur[], Int32(wp), W, X, Am, P, V))
  ccall((:az_samples_destroy, LIB), Int32, (Ptr{Cvoid},), cur[])
  return (; W, X, A = Am, P, V)
end

# ---- checkpoint interop (alphazero.jl_b200/checkpoint.py documents the format) --------------------------------------------
# Written next to the Serialization files of save_env (src/ui/session.jl:92-108) so the engine side can resume from a
# Julia session and vice versa.
function write_azb(path, gspec, nn)
  blob = flux_blob(Network.to_cpu(nn))
  header = JSON3.write((kind = nn isa NetLib.ResNet ? "resnet" : "simplenet", game = game_name(gspec),
                        hyperparams = Network.hyperparams(nn), num_params = length(blob), dtype = "float32", order = "flux"))
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
  blob = read_azb_blob(path)
  q = 0
  function fill!(dst)
    copyto!(dst, reshape(view(blob, (q + 1):(q + length(dst))), size(dst)))
    q += length(dst)
  end
  for chain in (nn.common, nn.vhead, nn.phead)
    for l in Flux.modules(chain)
      if l isa Flux.Conv
        fill!(l.weight); fill!(l.bias)
      elseif l isa Flux.BatchNorm
        fill!(l.γ); fill!(l.β); fill!(l.μ); fill!(l.σ²)
      elseif l isa Flux.Dense
        fill!(l.weight); fill!(l.bias)
      end
    end
  end
  q == length(blob) || error("blob size mismatch")
  return nn
end
function write_azs(path, gspec, samples)   # get_experience(env) -> mem.azs
  n, A = length(samples), GI.num_actions(gspec)
  g = game_name(gspec)
  SB = Int(ccall((:az_game_state_bytes, LIB), Int32, (Int32,), ccall((:az_game_lookup, LIB), Int32, (Cstring,), g)))
  states = Matrix{UInt8}(undef, SB, n); pi = zeros(Float64, A, n)
  for (i, s) in enumerate(samples)
    states[:, i] = state_to_bytes(gspec, s.s)
    pi[GI.actions_mask(GI.init(gspec, s.s)), i] = s.π
  end
  header = JSON3.write((game = g, num_samples = n, state_bytes = SB, num_actions = A))
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

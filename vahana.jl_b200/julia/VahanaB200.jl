# VahanaB200.jl — thin `ccall` wrapper over include/vahana_b200.h with Vahana's function names.
#
# UNTESTED in this repository: Julia is not installed in the build image (the ctypes mirror vahana.jl_b200/__init__.py
# is what the tests drive; this file follows it call by call).  It is the binding a Vahana maintainer would add where
# `create_model` generates per-model methods (src/Simulation.jl:115-207): the names, argument meaning and error behaviour
# are the reference's; what differs is that a transition is the *name* of a registered CUDA functor set instead of a closure.
module VahanaB200

export ModelTypes, register_agenttype!, register_edgetype!, register_param!, register_global!, create_model,
       create_simulation, copy_simulation, finish_simulation!, finish_init!, apply!, apply,
       AgentID, Edge, agent_id, type_nr, process_nr, agent_nr,
       add_agent!, add_agents!, add_edge!, add_edges!, remove_edges!,
       agentstate, agentstate_flexible, edges, neighborids, edgestates, neighborstates, neighborstates_flexible,
       neighborstates_iter, neighborstates_flexible_iter, checked,
       num_edges, has_edge, num_agents, all_agents, all_agentids, all_edges,
       param, set_param!, get_global, set_global!, push_global!, modify_global!,
       add_raster!, connect_raster_neighbors!, move_to!, cellid, calc_raster, calc_rasterstate, rastervalues, calc_raster_num_edges, raster_ids,
       random_pos, random_cell,
       enable_asserts

const LIB = get(ENV, "VAHANA_B200_LIB", joinpath(@__DIR__, "..", "csrc", "build", "libvahana_b200.so"))
const EDGE_REF = 256
const AgentID = UInt64

# ---- AgentID packing (src/Agent.jl:30-118): type:8 | rank:20 | nr:36, nr 1-based ----------------------------------
const SHIFT_TYPE = 56
const SHIFT_RANK = 36
agent_id(typeid, rank, nr) = (AgentID(typeid) << SHIFT_TYPE) + (AgentID(rank) << SHIFT_RANK) + AgentID(nr)
type_nr(id::AgentID) = Int(id >> SHIFT_TYPE)
process_nr(id::AgentID) = Int((id >> SHIFT_RANK) & 0xfffff)
agent_nr(id::AgentID) = Int(id & ((AgentID(1) << SHIFT_RANK) - 1))

"Edge{T}(from, state): src/Edge.jl:17-20"
struct Edge{T}
    from::AgentID
    state::T
end

# ---- C structs of include/vahana_b200.h ------------------------------------------------------------------------------
struct AgentTypeDesc
    name::Cstring
    size::UInt32
    hints::UInt32
end
struct EdgeTypeDesc
    name::Cstring
    size::UInt32
    hints::UInt32
    target::Int32
    size_hint::UInt64
end
struct ModelDesc
    name::Cstring
    n_agent_types::UInt32
    agent_types::Ptr{AgentTypeDesc}
    n_edge_types::UInt32
    edge_types::Ptr{EdgeTypeDesc}
    param_size::UInt32
end

const AGENT_HINTS = Dict(:Immortal => 1, :Independent => 2)
const EDGE_HINTS = Dict(:Stateless => 1, :IgnoreFrom => 2, :SingleEdge => 4, :SingleType => 8, :IgnoreSourceState => 16,
                        :NumEdgesOnly => 1 | 2, :HasEdgeOnly => 1 | 2 | 4)        # src/ModelTypes.jl:165-186
const OPS = Dict{Any,Int}(+ => 0, * => 1, min => 2, max => 3, & => 4, | => 5)
const DTS = Dict{DataType,Int}(Int64 => 0, Float64 => 1, Bool => 2, Int32 => 3, Float32 => 4, UInt8 => 5)
const METRICS = Dict(:chebyshev => 0, :euclidean => 1, :manhatten => 2)             # src/Raster.jl:83 (the reference's spelling)
const ACC_EDGES, ACC_NEIGHBORIDS, ACC_EDGESTATES, ACC_NUM_EDGES, ACC_HAS_EDGE = 0, 1, 2, 3, 4

const ASSERTS = Ref(true)
"enable_asserts(flag): src/Vahana.jl:42-61"
enable_asserts(flag::Bool) = (ASSERTS[] = flag)

last_error() = unsafe_string(ccall((:vb_last_error, LIB), Cstring, ()))
function check(rc::Integer)
    rc == 0 && return nothing
    rc == 1 && throw(AssertionError(last_error()))     # the reference's @assert / @mayassert
    rc == 2 && throw(ArgumentError(last_error()))
    error("vahana_b200 [$rc]: $(last_error())")
end

# ---- model definition (src/ModelTypes.jl:81-230) -----------------------------------------------------------------------
mutable struct ModelTypes
    agenttypes::Vector{DataType}
    agenthints::Vector{UInt32}
    edgetypes::Vector{DataType}
    edgehints::Vector{UInt32}
    edgetargets::Vector{Int32}
    edgesizes::Vector{UInt64}
    params::Vector{Pair{Symbol,Any}}
    globals::Vector{Pair{Symbol,Any}}
    ModelTypes() = new(DataType[], UInt32[], DataType[], UInt32[], Int32[], UInt64[], Pair{Symbol,Any}[], Pair{Symbol,Any}[])
end

function register_agenttype!(types::ModelTypes, ::Type{T}, hints...) where T
    @assert !(T in types.agenttypes) "Each type can be added only once"
    @assert isbitstype(T) "Agenttypes $T must be a bitstype"
    @assert length(types.agenttypes) < 255 "Can not add new type, maximal number of agent types is reached"
    h = UInt32(0)
    for hint in hints
        @assert haskey(AGENT_HINTS, hint) "$hint is not a valid agent hint"
        h |= UInt32(AGENT_HINTS[hint])
    end
    push!(types.agenttypes, T); push!(types.agenthints, h)
    types
end
register_agenttype!(t::Type, hints...) = types -> register_agenttype!(types, t, hints...)

function register_edgetype!(types::ModelTypes, ::Type{T}, hints...; target = nothing, size = 0) where T
    @assert !(T in types.edgetypes) "Each type can be added only once"
    @assert isbitstype(T) "Edgetypes $T must be a bitstype"
    h = UInt32(0)
    for hint in hints
        @assert haskey(EDGE_HINTS, hint) "$hint is not a valid edge hint"
        h |= UInt32(EDGE_HINTS[hint])
    end
    fieldcount(T) == 0 && (h |= UInt32(1))                       # detect_stateless (src/ModelTypes.jl:148-153)
    target !== nothing && (h |= UInt32(8))                       # a target implies :SingleType (:188-206)
    if (h & 8) != 0
        @assert target !== nothing "The :SingleType hint needs the keyword `target`"
    end
    if (h & 8) != 0 && (h & 4) != 0
        @assert (h & 3) == 3 ":SingleType and :SingleEdge can only be combined with :Stateless and :IgnoreFrom"
    end
    push!(types.edgetypes, T); push!(types.edgehints, h)
    push!(types.edgetargets, target === nothing ? Int32(0) : Int32(findfirst(==(target), types.agenttypes)))
    push!(types.edgesizes, UInt64(size))
    types
end
register_edgetype!(t::Type, hints...; kw...) = types -> register_edgetype!(types, t, hints...; kw...)

register_param!(types::ModelTypes, name::Symbol, default) = (push!(types.params, name => default); types)
register_param!(name::Symbol, default) = types -> register_param!(types, name, default)
register_global!(types::ModelTypes, name::Symbol, default) = (push!(types.globals, name => default); types)
register_global!(name::Symbol, default) = types -> register_global!(types, name, default)

struct Model
    types::ModelTypes
    name::String
end
create_model(types::ModelTypes, name::String) = Model(types, name)
create_model(name::String) = types -> create_model(types, name)

# ---- simulation ------------------------------------------------------------------------------------------------------
mutable struct Simulation
    handle::Ptr{Cvoid}
    model::Model
    params::Dict{Symbol,Any}        # the device copy is the values packed in registration order (all isbits)
    globals::Dict{Symbol,Any}
    rasters::Dict{Symbol,Tuple}
end
typeid(sim::Simulation, ::Type{T}) where T = Cint(findfirst(==(T), sim.model.types.agenttypes))
edgeidx(sim::Simulation, ::Type{T}) where T = Cint(findfirst(==(T), sim.model.types.edgetypes) - 1)
ref(sim::Simulation, T) = T in sim.model.types.agenttypes ? typeid(sim, T) : Cint(EDGE_REF + edgeidx(sim, T))
refs(sim::Simulation, ts) = Cint[ref(sim, T) for T in (ts isa Union{Tuple,AbstractVector} ? ts : (ts,))]
edgehints(sim::Simulation, T) = sim.model.types.edgehints[edgeidx(sim, T) + 1]
issingle(sim, T) = (edgehints(sim, T) & 4) != 0
isstateless(sim, T) = (edgehints(sim, T) & 1) != 0
ignorefrom(sim, T) = (edgehints(sim, T) & 2) != 0

function pack_params(types::ModelTypes, values::Dict{Symbol,Any})
    io = IOBuffer()
    for (name, _) in types.params
        v = values[name]
        al = sizeof(v)                                                # natural alignment of the isbits value, as in the C struct
        while al > 0 && position(io) % al != 0
            write(io, UInt8(0))
        end
        write(io, v)
    end
    take!(io)
end

"create_simulation(model, params = nothing, globals = nothing): src/Simulation.jl:261-313"
function create_simulation(model::Model, params = nothing, globals = nothing; device::Integer = 0)
    check(ccall((:vb_init, LIB), Cint, (Cint,), device))
    t = model.types
    pvals = Dict{Symbol,Any}(t.params)
    params !== nothing && for (k, v) in pairs(params)
        pvals[Symbol(k)] = v
    end
    gvals = Dict{Symbol,Any}(t.globals)
    globals !== nothing && for (k, v) in pairs(globals)
        gvals[Symbol(k)] = v
    end
    anames = [string(T) for T in t.agenttypes]
    enames = [string(T) for T in t.edgetypes]
    pbytes = pack_params(t, pvals)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve anames enames begin
        ads = [AgentTypeDesc(Base.unsafe_convert(Cstring, anames[i]), fieldcount(t.agenttypes[i]) == 0 ? 0 : sizeof(t.agenttypes[i]),
                             t.agenthints[i]) for i in eachindex(anames)]
        eds = [EdgeTypeDesc(Base.unsafe_convert(Cstring, enames[i]), fieldcount(t.edgetypes[i]) == 0 ? 0 : sizeof(t.edgetypes[i]),
                            t.edgehints[i], t.edgetargets[i], t.edgesizes[i]) for i in eachindex(enames)]
        GC.@preserve ads eds begin
            md = Ref(ModelDesc(Base.unsafe_convert(Cstring, model.name), length(ads), pointer(ads), length(eds), pointer(eds), length(pbytes)))
            check(ccall((:vb_sim_create, LIB), Cint, (Ref{ModelDesc}, Ptr{UInt8}, Ref{Ptr{Cvoid}}), md, pbytes, h))
        end
    end
    sim = Simulation(h[], model, pvals, gvals, Dict{Symbol,Tuple}())
    check(ccall((:vb_set_config, LIB), Cint, (Ptr{Cvoid}, Cint, Cint), sim.handle, ASSERTS[], 1))
    sim
end
"copy_simulation(sim): src/Simulation.jl:500-510"
function copy_simulation(sim::Simulation)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:vb_sim_copy, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), sim.handle, h))
    Simulation(h[], sim.model, copy(sim.params), deepcopy(sim.globals), copy(sim.rasters))
end
"finish_simulation!(sim): src/Simulation.jl:551-573"
function finish_simulation!(sim::Simulation)
    sim.handle != C_NULL && ccall((:vb_sim_destroy, LIB), Cint, (Ptr{Cvoid},), sim.handle)
    sim.handle = C_NULL
    sim.globals
end

param(sim::Simulation, name::Symbol) = sim.params[name]                                        # src/Simulation.jl:587
function set_param!(sim::Simulation, name::Symbol, value)                                      # src/Simulation.jl:602
    sim.params[name] = value
    b = pack_params(sim.model.types, sim.params)
    check(ccall((:vb_set_param, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, UInt32), sim.handle, b, length(b)))
    sim
end
get_global(sim::Simulation, name::Symbol) = sim.globals[name]                                  # src/Global.jl:10-74 (host side)
set_global!(sim::Simulation, name::Symbol, value) = (sim.globals[name] = value)
push_global!(sim::Simulation, name::Symbol, value) = (sim.globals[name] = vcat(sim.globals[name], [value]))
modify_global!(sim::Simulation, name::Symbol, f) = set_global!(sim, name, f(get_global(sim, name)))

# ---- init phase ---------------------------------------------------------------------------------------------------------
"add_agents!(sim, agents): src/AgentMethods.jl:65-89 in bulk; returns the ids"
function add_agents!(sim::Simulation, agents::Vector{T}) where T
    ids = Vector{AgentID}(undef, length(agents))
    check(ccall((:vb_add_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, UInt64, Ptr{AgentID}), sim.handle, typeid(sim, T),
                fieldcount(T) == 0 ? C_NULL : pointer(agents), length(agents), ids))
    ids
end
add_agent!(sim::Simulation, agent::T) where T = add_agents!(sim, T[agent])[1]

"add_edges!(sim, from, to, states): src/EdgeMethods.jl:388-523 in bulk"
function add_edges!(sim::Simulation, from::Vector{AgentID}, to::Vector{AgentID}, states::Vector{T}) where T
    GC.@preserve states check(ccall((:vb_add_edges, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{AgentID}, Ptr{AgentID}, Ptr{Cvoid}, UInt64), sim.handle,
                                    edgeidx(sim, T), from, to, fieldcount(T) == 0 ? C_NULL : pointer(states), length(to)))
    nothing
end
add_edge!(sim::Simulation, from::AgentID, to::AgentID, state::T) where T = add_edges!(sim, AgentID[from], AgentID[to], T[state])
add_edge!(sim::Simulation, to::AgentID, edge::Edge{T}) where T = add_edge!(sim, edge.from, to, edge.state)
"remove_edges!(sim, to, T) / remove_edges!(sim, from, to, T): src/EdgeMethods.jl:527-599 (outside of transitions)"
remove_edges!(sim::Simulation, to::AgentID, ::Type{T}) where T =
    check(ccall((:vb_remove_edges, LIB), Cint, (Ptr{Cvoid}, Cint, AgentID, AgentID), sim.handle, edgeidx(sim, T), 0, to))
remove_edges!(sim::Simulation, from::AgentID, to::AgentID, ::Type{T}) where T =
    check(ccall((:vb_remove_edges, LIB), Cint, (Ptr{Cvoid}, Cint, AgentID, AgentID), sim.handle, edgeidx(sim, T), from, to))

"finish_init!(sim; partition, return_idmapping, partition_algo, distribute): src/Simulation.jl:403-476.
One rank: `return_idmapping` gives the identity mapping.  Several ranks: this wrapper only supports `distribute = false` (every rank
adds its own block with ids of its own rank); the hand-out from rank 0 (distribute!, src/MPI.jl:11-84) is implemented in the Python
mirror (`Simulation.finish_init`, `plan_distribution`) and needs a host-side exchange (MPI.jl) here."
function finish_init!(sim::Simulation; partition = Dict{AgentID, Int}(), return_idmapping = false, partition_algo = :EqualAgentNumbers,
                      distribute = true)
    r, w = Ref{Cint}(0), Ref{Cint}(1)
    ccall((:vb_comm_rank, LIB), Cint, (Ptr{Cint}, Ptr{Cint}), r, w)
    @assert !(distribute && w[] > 1) "finish_init!(distribute = true) on several ranks: use the Python mirror or distribute = false"
    check(ccall((:vb_finish_init, LIB), Cint, (Ptr{Cvoid},), sim.handle))
    if distribute && return_idmapping
        idmapping = Dict{AgentID, AgentID}()
        for T in sim.model.types.agenttypes, id in all_agentids(sim, T)
            idmapping[id] = id
        end
        return idmapping
    end
    sim
end

# ---- transitions -------------------------------------------------------------------------------------------------------
"apply!(sim, transition, call, read, write; add_existing, with_edge, seed): src/Simulation.jl:720-821.
`transition` names a CUDA functor set registered with VB_REGISTER_TRANSITION (one functor per called agent type)."
function apply!(sim::Simulation, transition::String, call, read, write; add_existing = [], with_edge = nothing, seed = 0)
    c, r, w, a = refs(sim, call), refs(sim, read), refs(sim, write), refs(sim, add_existing)
    we = with_edge === nothing ? Cint(-1) : edgeidx(sim, with_edge)
    check(ccall((:vb_apply, LIB), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cint}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Cint, Cint, UInt64),
                sim.handle, transition, c, length(c), r, length(r), w, length(w), a, length(a), we, seed))
    sim
end

"apply(sim, transition, call, read, write; kwargs...): the non-mutating form, src/Simulation.jl:847-856"
function apply(sim::Simulation, transition::String, call, read, write; kwargs...)
    newsim = copy_simulation(sim)
    apply!(newsim, transition, call, read, write; kwargs...)
    newsim
end

# ---- queries -------------------------------------------------------------------------------------------------------------
function num_agents(sim::Simulation, ::Type{T}) where T                   # src/Agent.jl:324-343
    n = Ref{UInt64}(0)
    check(ccall((:vb_num_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{UInt64}), sim.handle, typeid(sim, T), n))
    Int(n[])
end
function _all(sim::Simulation, ::Type{T}) where T
    n = Ref{UInt64}(0)
    check(ccall((:vb_all_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{AgentID}, UInt64, Ref{UInt64}), sim.handle, typeid(sim, T), C_NULL, C_NULL, 0, n))
    states = Vector{T}(undef, n[]); ids = Vector{AgentID}(undef, n[])
    check(ccall((:vb_all_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{AgentID}, UInt64, Ref{UInt64}), sim.handle, typeid(sim, T),
                fieldcount(T) == 0 ? C_NULL : pointer(states), ids, n[], n))
    states, ids
end
all_agents(sim::Simulation, ::Type{T}) where T = _all(sim, T)[1]          # src/Agent.jl:234-285
all_agentids(sim::Simulation, ::Type{T}) where T = _all(sim, T)[2]        # src/Agent.jl:287-313
function agentstate(sim::Simulation, id::AgentID, ::Type{T}) where T      # src/AgentMethods.jl:91-154
    out = Ref{T}()
    check(ccall((:vb_agentstate, LIB), Cint, (Ptr{Cvoid}, AgentID, Cint, Ptr{Cvoid}), sim.handle, id, typeid(sim, T), out))
    out[]
end
agentstate_flexible(sim::Simulation, id::AgentID) = agentstate(sim, id, sim.model.types.agenttypes[type_nr(id)])

# one row of an edge container through the accessor `what` (availability rules of docs/src/performance.md:129-136 are enforced by the engine)
function _row(sim::Simulation, to::AgentID, ::Type{T}, what::Integer) where T
    n = Ref{Int64}(0)
    e = edgeidx(sim, T)
    check(ccall((:vb_edges_of, LIB), Cint, (Ptr{Cvoid}, Cint, AgentID, Cint, Ptr{AgentID}, Ptr{Cvoid}, UInt64, Ref{Int64}), sim.handle, e, to, what, C_NULL, C_NULL, 0, n))
    n[] < 0 && return nothing
    (what == ACC_NUM_EDGES || what == ACC_HAS_EDGE || (isstateless(sim, T) && ignorefrom(sim, T))) && return (AgentID[], T[], Int(n[]))
    from = Vector{AgentID}(undef, n[]); states = Vector{T}(undef, n[])
    check(ccall((:vb_edges_of, LIB), Cint, (Ptr{Cvoid}, Cint, AgentID, Cint, Ptr{AgentID}, Ptr{Cvoid}, UInt64, Ref{Int64}), sim.handle, e, to, what,
                from, fieldcount(T) == 0 ? C_NULL : pointer(states), n[], n))
    (from, states, Int(n[]))
end
"edges(sim, id, T): src/EdgeMethods.jl:704-715 — `nothing` when the agent has no edge of that type"
function edges(sim::Simulation, to::AgentID, ::Type{T}) where T
    r = _row(sim, to, T, ACC_EDGES)
    r === nothing && return nothing
    es = [Edge{T}(r[1][i], r[2][i]) for i in 1:r[3]]
    issingle(sim, T) ? es[1] : es
end
function neighborids(sim::Simulation, to::AgentID, ::Type{T}) where T    # :717-761
    r = _row(sim, to, T, ACC_NEIGHBORIDS)
    r === nothing && return nothing
    issingle(sim, T) ? r[1][1] : r[1]
end
function edgestates(sim::Simulation, to::AgentID, ::Type{T}) where T     # :804-848
    r = _row(sim, to, T, ACC_EDGESTATES)
    r === nothing && return nothing
    issingle(sim, T) ? r[2][1] : r[2]
end
function neighborstates(sim::Simulation, to::AgentID, ::Type{T}, ::Type{A}) where {T,A}   # :764-802
    ids = neighborids(sim, to, T)
    ids === nothing && return nothing
    ids isa AgentID ? agentstate(sim, ids, A) : [agentstate(sim, i, A) for i in ids]
end
function neighborstates_flexible(sim::Simulation, to::AgentID, ::Type{T}) where T         # src/Edge.jl:303-335
    ids = neighborids(sim, to, T)
    ids === nothing && return nothing
    ids isa AgentID ? agentstate_flexible(sim, ids) : [agentstate_flexible(sim, i) for i in ids]
end
function neighborstates_iter(sim::Simulation, to::AgentID, ::Type{T}, ::Type{A}) where {T,A}   # src/EdgeMethods.jl:783-803
    @assert !issingle(sim, T) "neighborstates_iter is not defined for edgetypes with the :SingleEdge or :IgnoreFrom hint"
    ids = neighborids(sim, to, T)
    ids === nothing ? nothing : (agentstate(sim, i, A) for i in ids)
end
function neighborstates_flexible_iter(sim::Simulation, to::AgentID, ::Type{T}) where T
    @assert !issingle(sim, T) "neighborstates_flexible_iter is not defined for edgetypes with the :SingleEdge or :IgnoreFrom hint"
    ids = neighborids(sim, to, T)
    ids === nothing ? nothing : (agentstate_flexible(sim, i) for i in ids)
end
"checked(f, g, itr): src/Helpers.jl:33-37"
checked(f, g, itr; kwargs...) = isnothing(itr) ? nothing : g(f, itr; kwargs...)
function num_edges(sim::Simulation, to::AgentID, ::Type{T}) where T      # :850-869
    r = _row(sim, to, T, ACC_NUM_EDGES)
    r === nothing ? 0 : r[3]
end
function has_edge(sim::Simulation, to::AgentID, ::Type{T}) where T       # :871-892
    r = _row(sim, to, T, ACC_HAS_EDGE)
    r !== nothing && r[3] >= 1
end
function num_edges(sim::Simulation, ::Type{T}; write = false) where T    # src/Edge.jl:373-389
    n = Ref{UInt64}(0)
    check(ccall((:vb_num_edges_total, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{UInt64}), sim.handle, edgeidx(sim, T), write, n))
    Int(n[])
end
"all_edges(sim, T): src/EdgeMethods.jl:1005-1027 — vector of (to, Edge) pairs"
function all_edges(sim::Simulation, ::Type{T}) where T
    n = Ref{UInt64}(0)
    e = edgeidx(sim, T)
    check(ccall((:vb_all_edges, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{AgentID}, Ptr{AgentID}, Ptr{Cvoid}, UInt64, Ref{UInt64}), sim.handle, e, C_NULL, C_NULL, C_NULL, 0, n))
    to = Vector{AgentID}(undef, n[]); from = Vector{AgentID}(undef, n[]); states = Vector{T}(undef, n[])
    check(ccall((:vb_all_edges, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{AgentID}, Ptr{AgentID}, Ptr{Cvoid}, UInt64, Ref{UInt64}), sim.handle, e, to, from,
                fieldcount(T) == 0 ? C_NULL : pointer(states), n[], n))
    [(to[i], Edge{T}(from[i], states[i])) for i in 1:Int(n[])]
end

# ---- reductions -------------------------------------------------------------------------------------------------------
"mapreduce(sim, field, op, T; init, equals): the reference's mapreduce(sim, a -> a.field, op, T) (src/AgentMethods.jl:533-565,
src/EdgeMethods.jl:972-994) for op in (+, *, min, max, &, |); `equals = v` maps a -> (a.field == v) first; `field = nothing` maps _ -> 1.
The map is a field selector because it has to run on the device; any other closure is registered as a map functor (next method)."
function Base.mapreduce(sim::Simulation, field::Union{Symbol,Nothing}, op, ::Type{T}; init = nothing, equals = nothing) where T
    if field === nothing
        off, fdt, FT = 0, -1, Int64
    else
        i = findfirst(==(field), fieldnames(T))
        FT = fieldtype(T, i)
        off, fdt = Int(fieldoffset(T, i)), DTS[FT]
    end
    RT = equals !== nothing ? ((op == (&) || op == (|)) ? Bool : Int64) :
         FT <: AbstractFloat ? Float64 : (FT == Bool && (op == (&) || op == (|)) ? Bool : Int64)
    _mapreduce(sim, ref(sim, T), off, fdt, equals, op, RT, init)
end
# the result type reaches ccall as a static type parameter (ccall's argument types cannot depend on local variables)
function _mapreduce(sim::Simulation, tref, off, fdt, equals, op, ::Type{RT}, init) where RT
    out = Ref{RT}()
    initref = init === nothing ? C_NULL : Ref{RT}(RT(init))
    GC.@preserve initref check(ccall((:vb_mapreduce, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Int64, Cint, Cint, Ptr{Cvoid}, Ref{RT}),
                                     sim.handle, tref, off, fdt, equals !== nothing, equals === nothing ? 0 : Int64(equals), OPS[op], DTS[RT],
                                     init === nothing ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, initref), out))
    out[]
end

"mapreduce(sim, mapname::String, op, T, R = Float64; init): the map is a functor registered with VB_REGISTER_MAP for T, i.e. any closure of
the reference's mapreduce(sim, f, op, T; datatype = R), e.g. b -> b.x - b.y (docs/examples/tutorial1.jl:548)."
function Base.mapreduce(sim::Simulation, mapname::String, op, ::Type{T}, ::Type{R} = Float64; init = nothing) where {T, R}
    out = Ref{R}()
    initref = init === nothing ? C_NULL : Ref{R}(R(init))
    GC.@preserve initref check(ccall((:vb_mapreduce_fn, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint, Cint, Cint, Ptr{Cvoid}, Ref{R}),
                                     sim.handle, mapname, ref(sim, T), OPS[op], DTS[R],
                                     init === nothing ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, initref), out))
    out[]
end

# ---- rasters (src/Raster.jl) -------------------------------------------------------------------------------------------
"add_raster!(sim, name, dims, agent_constructor): src/Raster.jl:32-54 — cells are agents in column-major order; returns the ids"
function add_raster!(sim::Simulation, name::Symbol, dims::NTuple{N,Int}, agent_constructor) where N
    cells = [agent_constructor(Tuple(ci)) for ci in CartesianIndices(dims)]
    T = eltype(cells)
    ids = Array{AgentID,N}(undef, dims)
    d = Int64[dims...]
    GC.@preserve cells check(ccall((:vb_add_raster, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint, Ptr{Int64}, Cint, Ptr{Cvoid}, Ptr{AgentID}), sim.handle, string(name), N, d,
                                   typeid(sim, T), fieldcount(T) == 0 ? C_NULL : pointer(cells), ids))
    sim.rasters[name] = (dims, T)
    ids
end
"connect_raster_neighbors!(sim, name, edge_constructor; distance, metric, periodic): src/Raster.jl:139-167"
function connect_raster_neighbors!(sim::Simulation, name::Symbol, edge_constructor; distance = 1, metric::Symbol = :chebyshev, periodic = true)
    st = edge_constructor((0,), (0,))            # one state for all edges (the docs' usage: (_, _) -> Neighbor())
    T = typeof(st)
    r = Ref(st)
    GC.@preserve r check(ccall((:vb_connect_raster_neighbors, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint, Cdouble, Cint, Cint, Ptr{Cvoid}), sim.handle, string(name),
                               edgeidx(sim, T), distance, METRICS[metric], periodic, fieldcount(T) == 0 ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, r)))
    nothing
end
"move_to!(sim, name, id, pos, edge_from_raster, edge_to_raster; distance, metric, periodic, only_surrounding): src/Raster.jl:437-477"
function move_to!(sim::Simulation, name::Symbol, id::AgentID, pos, edge_from_raster, edge_to_raster;
                  distance = 0, metric::Symbol = :chebyshev, periodic = true, only_surrounding = false)
    p = Int64[pos...]
    ef = edge_from_raster === nothing ? Cint(-1) : edgeidx(sim, typeof(edge_from_raster))
    et = edge_to_raster === nothing ? Cint(-1) : edgeidx(sim, typeof(edge_to_raster))
    rf = edge_from_raster === nothing || fieldcount(typeof(edge_from_raster)) == 0 ? nothing : Ref(edge_from_raster)
    rt = edge_to_raster === nothing || fieldcount(typeof(edge_to_raster)) == 0 ? nothing : Ref(edge_to_raster)
    GC.@preserve rf rt check(ccall((:vb_move_to, LIB), Cint, (Ptr{Cvoid}, Cstring, AgentID, Ptr{Int64}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cdouble, Cint, Cint, Cint),
                                   sim.handle, string(name), id, p, ef, rf === nothing ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, rf),
                                   et, rt === nothing ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, rt), distance, METRICS[metric], periodic, only_surrounding))
    nothing
end
"""broadcastids (src/MPI.jl:59-73, src/Raster.jl:64-75), receiver side: a raster over agents that exist already, `ids` in CartesianIndices
order; after finish_init!(distribute = true) some of them live on other ranks and the raster's read-outs join the ranks"""
function set_raster!(sim::Simulation, name::Symbol, ids::Array{AgentID,N}, ::Type{T}) where {N,T}
    d = Int64[size(ids)...]
    check(ccall((:vb_set_raster, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint, Ptr{Int64}, Cint, Ptr{AgentID}), sim.handle, string(name), N, d, typeid(sim, T), ids))
    sim.rasters[name] = (size(ids), T)
    nothing
end
"share of sampled edges whose source key passed the prefilter at the last check (negative: never checked); engine-side policy, no reference counterpart"
function last_pass_rate(sim::Simulation)
    r = Ref{Float64}(-1.0)
    check(ccall((:vb_last_pass_rate, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}), sim.handle, r))
    r[]
end
function cellid(sim::Simulation, name::Symbol, pos)                       # src/Raster.jl:403-405
    id = Ref{AgentID}(0)
    check(ccall((:vb_cellid, LIB), Cint, (Ptr{Cvoid}, Cstring, Ptr{Int64}, Ref{AgentID}), sim.handle, string(name), Int64[pos...], id))
    id[]
end
"random_pos(sim, name[, weights]) / random_cell(sim, name[, weights]): src/Raster.jl:512-577 (host side; uniform or weight-proportional)"
function random_pos(sim::Simulation, name::Symbol, weights::Union{Array,Nothing} = nothing)
    dims, _ = sim.rasters[name]
    positions = CartesianIndices(dims)
    weights === nothing && return positions[rand(1:length(positions))]
    @assert dims == size(weights) "`weights` must have the same dimension as the raster :$(name)"
    c = cumsum(vec(weights))
    positions[min(searchsortedfirst(c, rand() * c[end]), length(c))]
end
random_cell(sim::Simulation, name::Symbol, weights::Union{Array,Nothing} = nothing) = cellid(sim, name, Tuple(random_pos(sim, name, weights)))

"rastervalues(sim, name, field) / calc_rasterstate(sim, name, field): src/Raster.jl:290-387 with f = c -> c.field"
function rastervalues(sim::Simulation, name::Symbol, field::Symbol)
    dims, T = sim.rasters[name]
    i = findfirst(==(field), fieldnames(T)); FT = fieldtype(T, i)
    out = Array{FT}(undef, dims)
    check(ccall((:vb_rastervalues, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint, Cint, Ptr{Cvoid}), sim.handle, string(name), fieldoffset(T, i), DTS[FT], out))
    out
end
calc_rasterstate(sim::Simulation, name::Symbol, field::Symbol) = rastervalues(sim, name, field)
"calc_raster(sim, name, id -> num_edges(sim, id, E), Int64, [E]): src/Raster.jl:206-236 for the edge-count read-out"
function calc_raster_num_edges(sim::Simulation, name::Symbol, ::Type{E}) where E
    dims, _ = sim.rasters[name]
    out = Array{Int64}(undef, dims)
    check(ccall((:vb_calc_raster_num_edges, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint, Ptr{Int64}), sim.handle, string(name), edgeidx(sim, E), out))
    out
end

"raster_ids(sim, name): sim.rasters[name] of the reference, the grid of cell ids (on several ranks the grid handed out by finish_init!)"
function raster_ids(sim::Simulation, name::Symbol)
    dims = sim.rasters[name][1]
    ids = Array{AgentID}(undef, dims)
    nd = Ref{Cint}(0); d = zeros(Int64, 4)
    check(ccall((:vb_raster_info, LIB), Cint, (Ptr{Cvoid}, Cstring, Ref{Cint}, Ptr{Int64}, Ptr{AgentID}), sim.handle, string(name), nd, d, ids))
    ids
end
"calc_raster(sim, raster, f, f_returns, accessible): src/Raster.jl:206-236, the general form as a host loop over the cells (one rank)"
function calc_raster(sim::Simulation, name::Symbol, f, ::Type{R}, accessible = DataType[]) where R
    ids = raster_ids(sim, name)
    out = zeros(R, size(ids))
    for (idx, id) in enumerate(ids)
        out[idx] = f(id)
    end
    out
end

"calc_rasterstate(sim, raster, mapname::String, f_returns): src/Raster.jl:238-280 with f a map functor registered (VB_REGISTER_MAP) for the cells' type"
function calc_rasterstate(sim::Simulation, name::Symbol, mapname::String, ::Type{R} = Float64) where R
    dims = sim.rasters[name][1]
    isf = R <: AbstractFloat
    out = isf ? Vector{Float64}(undef, prod(dims)) : Vector{Int64}(undef, prod(dims))
    check(ccall((:vb_calc_rasterstate_fn, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring, Cint, Ptr{Cvoid}), sim.handle, String(name), mapname, isf, out))
    reshape(R.(out), dims)
end

end # module

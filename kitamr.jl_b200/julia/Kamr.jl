# Kamr.jl — the host shim that binds libkamr.so (include/kamr.h) into KitAMR.jl.
#
# Drop this file into the reference as `src/GPU/Kamr.jl` and `include("GPU/Kamr.jl")` it at the end of
# `src/KitAMR.jl`.  It flattens `ka` after every `amr_recover!` (src/Solver/AMR.jl:54), uploads the flat arrays and
# replaces the body of the time-march loop (`slope!` → `flux!` → `iterate!`, src/Solver/Solver.jl:65-67) by `ccall`s.
# Everything else of KitAMR (configuration, p4est mesh, adaptation, partition, IO) is untouched.
#
# STATUS: written against the reference source, NEVER EXECUTED — the build image has no `julia`, no libp4est and no MPI.
# The same arrays, in the same order, are produced for the synthetic forest by `kitamr.jl_b200/model.py::build_rank_view`,
# and that path is what the parity tests exercise.  `dump_reference_step` at the bottom writes the fixture that pins the
# CPU oracle against the real reference the first time somebody runs this under Julia (SURVEY.md §8c: "parity unpinned").
module Kamr

using ..KitAMR
using ..KitAMR: KA, PsData, AbstractPsData, InsideSolidData, GhostPsData, GhostInsideSolidData, SolidNeighbor,
                FullFace, HangingFace, BackHangingFace, DomainFace, CAIDVM, DVM, CAIDVM_Marching, CIP_Marching, Euler,
                Maxwellian, SuperSonicInflow, UniformOutflow, InterpolatedOutflow, get_bc, min_cell_level, cell_level,
                residual_comm!, _ghost_comm_arrays, pw_mirror_quadrant, PointerWrapper, P4estPsData
using MPI

const LIB = get(ENV, "KAMR_LIB", "libkamr.so")

# ---------------------------------------------------------------------------------------------------- C structs
struct Config                         # == kamr_config (include/kamr.h)
    dim::Int32; ndf::Int32; flux_type::Int32; marching::Int32
    K::Float64; Pr::Float64; gamma::Float64; omega::Float64; mu_ref::Float64
    device::Int32; rank::Int32; nranks::Int32
    stream::Ptr{Cvoid}
end

struct CIb                            # == kamr_ib
    n_solid::Int32
    solid_cell::Ptr{Int32}; solid_nb_off::Ptr{Int32}; solid_nb_ids::Ptr{Int32}
    n_sn::Int32
    sn_donor::Ptr{Int32}; sn_solid::Ptr{Int32}; sn_faceid::Ptr{Int32}
    sn_aux::Ptr{Float64}; sn_normal::Ptr{Float64}; sn_bc::Ptr{Float64}
    sn_nb_off::Ptr{Int32}; sn_nb_ids::Ptr{Int32}
    cvc_off::Ptr{Int32}; cvc_index::Ptr{Int32}; cvc_gas_w::Ptr{Float64}; cvc_solid_w::Ptr{Float64}
end

struct CMesh                          # == kamr_mesh; field order and types as in include/kamr.h
    n_local::Int32; n_ghost::Int32; n_solidnbr::Int32
    ds::Ptr{Float64}; mid::Ptr{Float64}; bound_enc::Ptr{Int32}; ps_level::Ptr{Int32}; cell_grid::Ptr{Int32}
    n_grid::Int32; grid_off::Ptr{Int64}; v_level::Ptr{Int8}; v_weight::Ptr{Float64}; v_mid::Ptr{Float64}
    nb_state::Ptr{Int32}; nb_off::Ptr{Int32}; nb_ids::Ptr{Int32}
    ps_maxlevel::Int32; ps_minlevel::Int32
    n_face::Int32; face_kind::Ptr{Int32}; face_here::Ptr{Int32}; face_there::Ptr{Int32}; face_dir::Ptr{Int32}
    face_rot::Ptr{Float64}; face_mid::Ptr{Float64}; face_there_mid::Ptr{Float64}
    n_bc::Int32; bc_type::Ptr{Int32}; bc_prim::Ptr{Float64}
    n_peer::Int32; peer_rank::Ptr{Int32}; send_off::Ptr{Int32}; send_cells::Ptr{Int32}; recv_off::Ptr{Int32}
    ib::Ptr{CIb}
end

# KAMR_FACE_*, KAMR_BC_*, KAMR_DL_* of include/kamr.h
const FACE_DOMAIN, FACE_FULL, FACE_HANGING, FACE_BACKHANGING = Int32(0), Int32(1), Int32(2), Int32(3)
const DL_DF, DL_SDF, DL_FLUX, DL_W, DL_PRIM, DL_QF, DL_SW, DL_MFLUX =
    UInt32(1), UInt32(2), UInt32(4), UInt32(8), UInt32(16), UInt32(32), UInt32(64), UInt32(128)
bc_code(::Type{Maxwellian}) = Int32(0)
bc_code(::Type{SuperSonicInflow}) = Int32(1)
bc_code(::Type{UniformOutflow}) = Int32(2)
bc_code(::Type{InterpolatedOutflow}) = Int32(3)

# ---------------------------------------------------------------------------------------------------- flat host model
"""
The flat arrays that cross the C-ABI (the Julia twin of `kitamr.jl_b200/model.py::HostMesh` / `HostState`).  All
vectors are owned by this object; `CMesh(flat)` only takes pointers, so calls are wrapped in `GC.@preserve flat`.
Cell ids are 0-based: `[0,n_local)` local cells that own velocity data in p4est order, `[n_local, n_local+n_ghost)`
ghost cells in ghost-id order, then one pseudo-cell per `SolidNeighbor`.
"""
mutable struct Flat
    dim::Int; ndf::Int
    n_local::Int; n_ghost::Int; n_solidnbr::Int
    cells::Vector{Any}                 # the objects behind the cell ids (PsData | GhostPsData | SolidNeighbor)
    ds::Vector{Float64}; mid::Vector{Float64}; bound_enc::Vector{Int32}; ps_level::Vector{Int32}
    cell_grid::Vector{Int32}; grid_off::Vector{Int64}
    v_level::Vector{Int8}; v_weight::Vector{Float64}; v_mid::Vector{Float64}
    nb_state::Vector{Int32}; nb_off::Vector{Int32}; nb_ids::Vector{Int32}
    ps_maxlevel::Int; ps_minlevel::Int
    face_kind::Vector{Int32}; face_here::Vector{Int32}; face_there::Vector{Int32}; face_dir::Vector{Int32}
    face_rot::Vector{Float64}; face_mid::Vector{Float64}; face_there_mid::Vector{Float64}
    bc_type::Vector{Int32}; bc_prim::Vector{Float64}
    peer_rank::Vector{Int32}; send_off::Vector{Int32}; send_cells::Vector{Int32}; recv_off::Vector{Int32}
    # immersed boundary
    solid_cell::Vector{Int32}; solid_nb_off::Vector{Int32}; solid_nb_ids::Vector{Int32}
    sn_donor::Vector{Int32}; sn_solid::Vector{Int32}; sn_faceid::Vector{Int32}
    sn_aux::Vector{Float64}; sn_normal::Vector{Float64}; sn_bc::Vector{Float64}
    sn_nb_off::Vector{Int32}; sn_nb_ids::Vector{Int32}
    cvc_off::Vector{Int32}; cvc_index::Vector{Int32}; cvc_gas_w::Vector{Float64}; cvc_solid_w::Vector{Float64}
    cib::Base.RefValue{CIb}
    # state (per-cell column-major blocks, exactly VsData.df etc. back to back)
    vs_off::Vector{Int64}
    df::Vector{Float64}; sdf::Vector{Float64}; flux::Vector{Float64}
    w::Vector{Float64}; prim::Vector{Float64}; mflux::Vector{Float64}; qf::Vector{Float64}; sw::Vector{Float64}
    Flat() = new()
end

has_vs(x) = !(x isa InsideSolidData) && !(x isa GhostInsideSolidData)

"""
    flatten(p4est, ka) -> Flat

Walks `ka` once, in the reference's own orders:
* local cells: `ka.kdata.field.trees.data[tree][j]` (p4est quadrant order), skipping `InsideSolidData`
  (Solver/Initialize.jl:358-376);
* ghost cells: `ka.kdata.ghost.ghost_wrap` (ghost-id order, contiguous per source rank, Parallel/Ghost.jl:706-749);
* SolidNeighbor pseudo-cells: donor cells in local order, faces 1..2·DIM (Boundary/Immersed_boundary.jl:282-305);
* faces: `ka.kdata.field.faces` in order, Hanging / BackHanging faces expanded to one record per `FluxData`
  (Flux/Flux.jl:37-59);
* halo: the p4est ghost layer's `mirror_proc_offsets / mirror_proc_mirrors / proc_offsets` (Parallel/Ghost.jl:133-145).
Velocity grids are deduplicated by content so that the device stores identical grids once (the library would also
recognise duplicates itself).
"""
function flatten(p4est, ka::KA{DIM,NDF}) where {DIM,NDF}
    f = Flat(); f.dim = DIM; f.ndf = NDF
    M = DIM + 2
    trees = ka.kdata.field.trees.data
    ghosts = ka.kdata.ghost.ghost_wrap
    id = IdDict{Any,Int32}()                       # object -> 0-based cell id
    cells = Any[]
    for tree in trees, ps in tree
        has_vs(ps) || continue
        id[ps] = length(cells); push!(cells, ps)
    end
    f.n_local = length(cells)
    # ghost cells keep their ghost-id position even when they are InsideSolid placeholders (no velocity data): those get
    # a one-point dummy grid, nothing on the path reads them
    for g in ghosts
        id[g] = length(cells); push!(cells, g)
    end
    f.n_ghost = length(ghosts)
    sn_list = SolidNeighbor{DIM,NDF}[]
    for i in 1:f.n_local
        ps = cells[i]
        for fid in 1:2*DIM
            nb = ps.neighbor.data[fid]
            (isempty(nb) || nb[1] === nothing) && continue
            if nb[1] isa SolidNeighbor
                id[nb[1]] = length(cells); push!(cells, nb[1]); push!(sn_list, nb[1])
            end
        end
    end
    f.n_solidnbr = length(sn_list)
    f.cells = cells
    nc = length(cells)

    # ---- per-cell geometry
    f.ds = zeros(nc * DIM); f.mid = zeros(nc * DIM)
    f.bound_enc = zeros(Int32, nc); f.ps_level = zeros(Int32, nc)
    for (c, x) in enumerate(cells)
        f.ds[(c-1)*DIM+1:c*DIM] .= x.ds
        f.mid[(c-1)*DIM+1:c*DIM] .= x.midpoint
        # SolidNeighbor and solid ghost cells: < 0; donors: > 0 (PsData.bound_enc)
        f.bound_enc[c] = x isa SolidNeighbor ? Int32(-abs(x.bound_enc)) : Int32(x.bound_enc)
        f.ps_level[c] = x isa SolidNeighbor ? cell_level(ka, x.solid_cell) : cell_level(ka, x)
    end

    # ---- velocity grids, deduplicated by content
    grid_of = Dict{UInt,Int32}()
    f.cell_grid = zeros(Int32, nc); f.grid_off = Int64[0]
    f.v_level = Int8[]; f.v_weight = Float64[]; f.v_mid = Float64[]
    dummy = Int32(-1)
    for (c, x) in enumerate(cells)
        if !has_vs(x)
            if dummy < 0
                dummy = Int32(length(f.grid_off) - 1)
                push!(f.v_level, Int8(0)); push!(f.v_weight, 1.0); append!(f.v_mid, ones(DIM))
                push!(f.grid_off, f.grid_off[end] + 1)
            end
            f.cell_grid[c] = dummy; continue
        end
        vs = x.vs_data
        key = hash(vs.level, hash(vs.midpoint))
        g = get(grid_of, key, Int32(-1))
        if g < 0
            g = Int32(length(f.grid_off) - 1); grid_of[key] = g
            append!(f.v_level, vs.level); append!(f.v_weight, vs.weight)
            append!(f.v_mid, vec(vs.midpoint))          # column-major [n × DIM] == planes of n points
            push!(f.grid_off, f.grid_off[end] + vs.vs_num)
        end
        f.cell_grid[c] = g
    end

    # a periodic alias is a shallow copy of the real neighbour with a shifted midpoint (Boundary/Period.jl:1-12): it
    # shares `vs_data` with its owner, so it is resolved through that
    vs_owner = IdDict{Any,Int32}()
    for (c, x) in enumerate(cells)
        has_vs(x) && !(x isa SolidNeighbor) && (vs_owner[x.vs_data] = Int32(c - 1))
    end
    cid(x) = haskey(id, x) ? id[x] : (has_vs(x) ? get(vs_owner, x.vs_data, Int32(-1)) : Int32(-1))
    # ---- slope neighbours of local cells: PsData.neighbor (Mesh/Neighbor.jl:31-63), list order preserved
    f.nb_state = zeros(Int32, f.n_local * 2 * DIM); f.nb_off = Int32[0]; f.nb_ids = Int32[]
    for i in 1:f.n_local
        ps = cells[i]
        for fid in 1:2*DIM
            f.nb_state[(i-1)*2*DIM+fid] = ps.neighbor.state[fid]
            for nb in ps.neighbor.data[fid]
                (nb === nothing || cid(nb) < 0) && continue        # boundary / InsideSolidData: never read
                push!(f.nb_ids, cid(nb))
            end
            push!(f.nb_off, length(f.nb_ids))
        end
    end
    f.ps_maxlevel = ka.kinfo.config.solver.AMR_PS_MAXLEVEL
    f.ps_minlevel = min_cell_level(ka)             # already all-reduced, Flux/Slope.jl:953-967

    # ---- domain boundary conditions: one record per DomainFace (function-valued bc evaluated at the face midpoint)
    f.bc_type = Int32[]; f.bc_prim = Float64[]
    f.face_kind = Int32[]; f.face_here = Int32[]; f.face_there = Int32[]; f.face_dir = Int32[]
    f.face_rot = Float64[]; f.face_mid = Float64[]; f.face_there_mid = Float64[]
    function emit(kind, here, there, dir, rot, fmid, tmid)
        push!(f.face_kind, kind); push!(f.face_here, here); push!(f.face_there, there); push!(f.face_dir, dir - 1)
        push!(f.face_rot, rot); append!(f.face_mid, fmid); append!(f.face_there_mid, tmid)
    end
    for face in ka.kdata.field.faces
        if face isa DomainFace
            push!(f.bc_type, bc_code(typeof(face).parameters[3]))
            append!(f.bc_prim, get_bc(face.domain.bc; intersect_point = face.midpoint))
            emit(FACE_DOMAIN, id[face.ps_data], Int32(length(f.bc_type) - 1), face.direction, face.rot, face.midpoint,
                 face.midpoint)
        elseif face isa FullFace
            cid(face.there_data) >= 0 || continue
            emit(FACE_FULL, id[face.here_data], cid(face.there_data), face.direction, face.rot, face.midpoint,
                 face.there_data.midpoint)          # periodic aliases carry the shifted midpoint (Boundary/Period.jl:1-12)
        elseif face isa HangingFace
            for (k, there) in enumerate(face.there_data)
                cid(there) >= 0 || continue
                emit(FACE_HANGING, id[face.here_data], cid(there), face.direction, face.rot, face.midpoint[k], there.midpoint)
            end
        elseif face isa BackHangingFace
            for (k, here) in enumerate(face.here_data)
                emit(FACE_BACKHANGING, id[here], cid(face.there_data), face.direction, face.rot, face.midpoint[k],
                     face.there_data.midpoint)
            end
        end
    end
    # ---- halo: mirrors per destination rank / ghosts per source rank
    f.peer_rank = Int32[]; f.send_off = Int32[0]; f.send_cells = Int32[]; f.recv_off = Int32[0]
    if MPI.Comm_size(MPI.COMM_WORLD) > 1
        mpisize, proc_offsets, mpo, mpm = _ghost_comm_arrays(ka.kinfo.forest.ghost)
        gp = PointerWrapper(ka.kinfo.forest.ghost); pp = PointerWrapper(p4est)
        for r in 0:mpisize-1
            ns = mpo[r+2] - mpo[r+1]; nr = proc_offsets[r+2] - proc_offsets[r+1]
            (ns == 0 && nr == 0) && continue
            push!(f.peer_rank, r)
            for k in mpo[r+1]+1:mpo[r+2]
                pq = pw_mirror_quadrant(pp, gp, mpm[k] + 1)
                dp = PointerWrapper(P4estPsData, pq.p.user_data[])
                ps = unsafe_pointer_to_objref(pointer(dp.ps_data))
                has_vs(ps) && push!(f.send_cells, id[ps])    # InsideSolid mirrors carry no data; the peer lists no ghost
            end
            push!(f.send_off, length(f.send_cells))
            push!(f.recv_off, proc_offsets[r+2])
        end
    end

    # ---- immersed boundary tables (Boundary/Immersed_boundary.jl:207-213, 282-305, 366-373, 226-277)
    f.solid_cell = Int32[]; f.solid_nb_off = Int32[0]; f.solid_nb_ids = Int32[]
    f.sn_donor = Int32[]; f.sn_solid = Int32[]; f.sn_faceid = Int32[]
    f.sn_aux = Float64[]; f.sn_normal = Float64[]; f.sn_bc = Float64[]
    f.sn_nb_off = Int32[0]; f.sn_nb_ids = Int32[]
    f.cvc_off = Int32[0]; f.cvc_index = Int32[]; f.cvc_gas_w = Float64[]; f.cvc_solid_w = Float64[]
    for ib in ka.kdata.field.immersed_boundaries
        for sc in ib.solid_cells
            (haskey(id, sc) && id[sc] < f.n_local) || continue
            fluid = [nb[1] for nb in sc.neighbor.data        # first element of every face AND corner list, :207-213
                     if !isempty(nb) && nb[1] !== nothing && haskey(id, nb[1]) && nb[1].bound_enc >= 0]
            isempty(fluid) && continue
            push!(f.solid_cell, id[sc])
            append!(f.solid_nb_ids, Int32[id[x] for x in fluid]); push!(f.solid_nb_off, length(f.solid_nb_ids))
        end
    end
    for sn in sn_list
        donor = first(c for c in cells[1:f.n_local] if any(d -> !isempty(d) && d[1] === sn, c.neighbor.data[1:2*DIM]))
        push!(f.sn_donor, id[donor]); push!(f.sn_solid, id[sn.solid_cell]); push!(f.sn_faceid, sn.faceid - 1)
        append!(f.sn_aux, sn.aux_point); append!(f.sn_normal, sn.normal)
        ibobj = ka.kinfo.config.IB[donor.bound_enc]
        append!(f.sn_bc, get_bc(ibobj.bc; intersect_point = sn.aux_point, ib = ibobj))
        for fid in 1:2*DIM                                       # image_df's fluid cells, :366-373 (the donor goes last,
            nb = donor.neighbor.data[fid]                        #  the library appends it)
            (isempty(nb) || nb[1] === nothing || !haskey(id, nb[1]) || nb[1].bound_enc < 0) && continue
            push!(f.sn_nb_ids, id[nb[1]])
        end
        push!(f.sn_nb_off, length(f.sn_nb_ids))
        order = sortperm(sn.cvc.indices)
        append!(f.cvc_index, Int32.(sn.cvc.indices[order] .- 1))
        append!(f.cvc_gas_w, sn.cvc.gas_weights[order]); append!(f.cvc_solid_w, sn.cvc.solid_weights[order])
        push!(f.cvc_off, length(f.cvc_index))
    end

    # ---- state
    f.vs_off = Int64[0]
    for x in cells
        push!(f.vs_off, f.vs_off[end] + (f.grid_off[f.cell_grid[length(f.vs_off)]+2] - f.grid_off[f.cell_grid[length(f.vs_off)]+1]))
    end
    npts = f.vs_off[end]
    f.df = zeros(npts * NDF); f.sdf = zeros(npts * NDF * DIM); f.flux = zeros(npts * NDF)
    f.w = zeros(nc * M); f.prim = zeros(nc * M); f.mflux = zeros(nc * M); f.qf = zeros(nc * DIM); f.sw = zeros(nc * M * DIM)
    gather_state!(f)
    return f
end

"`ka` -> flat state (df of every cell with velocity data, w / prim of the local cells)"
function gather_state!(f::Flat)
    M = f.dim + 2
    for (c, x) in enumerate(f.cells)
        has_vs(x) || continue
        n = x.vs_data.vs_num
        o = f.vs_off[c] * f.ndf
        copyto!(f.df, o + 1, x.vs_data.df, 1, n * f.ndf)        # column-major [n × NDF] block == the ABI layout
        # the raw slopes are state too: on meshes with non-dyadic cell sizes the next sweep projects some finer
        # neighbours' slopes of the previous step (DESIGN.md §5); [n × NDF × DIM] column-major == the ABI layout
        c <= f.n_local && copyto!(f.sdf, f.vs_off[c] * f.ndf * f.dim + 1, x.vs_data.sdf, 1, n * f.ndf * f.dim)
        if c <= f.n_local
            f.w[(c-1)*M+1:c*M] .= x.w; f.prim[(c-1)*M+1:c*M] .= x.prim
        end
    end
    return f
end

"flat state -> `ka` (what the event about to run reads: SURVEY.md Appendix D); `mask` = the KAMR_DL_* bits downloaded"
function scatter_state!(f::Flat, mask::UInt32)
    D, K, M = f.dim, f.ndf, f.dim + 2
    for c in 1:f.n_local
        x = f.cells[c]; n = x.vs_data.vs_num; o = f.vs_off[c]
        mask & DL_DF != 0 && copyto!(x.vs_data.df, 1, f.df, o * K + 1, n * K)
        mask & DL_SDF != 0 && copyto!(x.vs_data.sdf, 1, f.sdf, o * K * D + 1, n * K * D)
        mask & DL_W != 0 && (x.w .= @view f.w[(c-1)*M+1:c*M])
        mask & DL_PRIM != 0 && (x.prim .= @view f.prim[(c-1)*M+1:c*M])
        mask & DL_QF != 0 && (x.qf .= @view f.qf[(c-1)*D+1:c*D])
        mask & DL_SW != 0 && (vec(x.sw) .= @view f.sw[(c-1)*M*D+1:c*M*D])
    end
    if mask & DL_SW != 0                                           # ghost sw for the Löhner sensor (Criteria.jl:103-169)
        for c in f.n_local+1:f.n_local+f.n_ghost
            x = f.cells[c]; has_vs(x) || continue
            vec(x.sw) .= @view f.sw[(c-1)*M*D+1:c*M*D]
        end
    end
    return nothing
end

function CMesh(f::Flat)
    p(v) = pointer(v)
    f.cib = Ref(CIb(length(f.solid_cell), p(f.solid_cell), p(f.solid_nb_off), p(f.solid_nb_ids), f.n_solidnbr,
                    p(f.sn_donor), p(f.sn_solid), p(f.sn_faceid), p(f.sn_aux), p(f.sn_normal), p(f.sn_bc),
                    p(f.sn_nb_off), p(f.sn_nb_ids), p(f.cvc_off), p(f.cvc_index), p(f.cvc_gas_w), p(f.cvc_solid_w)))
    has_ib = f.n_solidnbr > 0 || !isempty(f.solid_cell)
    return CMesh(f.n_local, f.n_ghost, f.n_solidnbr, p(f.ds), p(f.mid), p(f.bound_enc), p(f.ps_level), p(f.cell_grid),
                 length(f.grid_off) - 1, p(f.grid_off), p(f.v_level), p(f.v_weight), p(f.v_mid),
                 p(f.nb_state), p(f.nb_off), p(f.nb_ids), f.ps_maxlevel, f.ps_minlevel,
                 length(f.face_kind), p(f.face_kind), p(f.face_here), p(f.face_there), p(f.face_dir),
                 p(f.face_rot), p(f.face_mid), p(f.face_there_mid),
                 length(f.bc_type), p(f.bc_type), p(f.bc_prim),
                 length(f.peer_rank), p(f.peer_rank), p(f.send_off), p(f.send_cells), p(f.recv_off),
                 has_ib ? Base.unsafe_convert(Ptr{CIb}, f.cib) : Ptr{CIb}(C_NULL))
end

# ---------------------------------------------------------------------------------------------------- the binding
mutable struct Context
    h::Ptr{Cvoid}
    flat::Union{Flat,Nothing}
end

last_error(h) = unsafe_string(ccall((:kamr_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(ctx::Context, rc) = rc == 0 || error("libkamr: " * last_error(ctx.h))   # the reference's error(...) behaviour

flux_code(::Type{CAIDVM}) = Int32(0)
flux_code(::Type{DVM}) = Int32(1)
flux_code(T) = error("libkamr builds the CAIDVM and DVM fluxes only (got $T)")
march_code(::Type{CAIDVM_Marching}) = Int32(0)
march_code(::Type{CIP_Marching}) = Int32(1)
march_code(::Type{Euler}) = Int32(2)
march_code(T) = error("libkamr builds CAIDVM_Marching, CIP_Marching and Euler only (got $T)")

"after `initialize` / `restart`: one context per MPI rank = per GPU"
function create(ka::KA{DIM,NDF}; device = MPI.Comm_rank(MPI.COMM_WORLD) % 8) where {DIM,NDF}
    s = ka.kinfo.config.solver; g = ka.kinfo.config.gas
    rank, nranks = MPI.Comm_rank(MPI.COMM_WORLD), MPI.Comm_size(MPI.COMM_WORLD)
    cfg = Config(DIM, NDF, flux_code(s.flux), march_code(s.time_marching), g.K, g.Pr, g.γ, g.ω, g.μᵣ,
                 device, rank, nranks, C_NULL)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:kamr_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error("libkamr: " * last_error(C_NULL))
    ctx = Context(h[], nothing)
    if nranks > 1                                   # the ncclUniqueId travels over MPI, as every KitAMR collective does
        uid = zeros(UInt8, 128)
        rank == 0 && ccall((:kamr_comm_unique_id, LIB), Cint, (Ptr{UInt8},), uid)
        MPI.Bcast!(uid, 0, MPI.COMM_WORLD)
        check(ctx, ccall((:kamr_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx.h, uid))
    end
    return ctx
end

destroy(ctx::Context) = (ccall((:kamr_destroy, LIB), Cint, (Ptr{Cvoid},), ctx.h); ctx.h = C_NULL; nothing)

"after initialize / restart / every `amr_recover!` (Solver/AMR.jl:54) — collective over the ranks, like amr_recover!"
function reflatten!(ctx::Context, p4est, ka::KA)
    flat = flatten(p4est, ka)
    GC.@preserve flat begin
        m = CMesh(flat)
        check(ctx, ccall((:kamr_upload_topology, LIB), Cint, (Ptr{Cvoid}, Ref{CMesh}), ctx.h, m))
        check(ctx, ccall((:kamr_upload_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                         ctx.h, flat.df, flat.w, flat.prim))
        check(ctx, ccall((:kamr_upload_aux, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                         ctx.h, flat.sdf, C_NULL, C_NULL))             # VsData.sdf as the adapt event left it
        check(ctx, ccall((:kamr_exchange_df, LIB), Cint, (Ptr{Cvoid},), ctx.h))
    end
    ctx.flat = flat
    return nothing
end

"residual_comm! (Solver/Finalize.jl:12-24) on the 2(DIM+2) local sums the device returned"
function residual_comm_from!(ka::KA{DIM}, res::Vector{Float64}) where {DIM}
    M = DIM + 2
    sumRes = res[1:M]; sumAvg = res[M+1:2M]
    MPI.Allreduce!(sumRes, +, MPI.COMM_WORLD); MPI.Allreduce!(sumAvg, +, MPI.COMM_WORLD)
    r = ka.kinfo.status.residual
    r.residual .= sqrt.(sumRes .* MPI.Comm_size(MPI.COMM_WORLD)) ./ (sumAvg .+ eps())   # as residual_comm! scales them
    return nothing
end

"""
Drop-in for the three calls of the time-march loop (Solver/Solver.jl:65-67); `adaptive_mesh_refinement!`, `limit_Δt!`,
`check!`, `check_for_convergence` stay the reference's.  Fused: the face flux never leaves the SM.
"""
function step!(ctx::Context, ka::KA{DIM}) where {DIM}
    st = ka.kinfo.status
    want = st.step % ka.kinfo.config.solver.ST_CHECK_INTERVAL == 0     # residual gating, Finalize.jl:5-11
    res = zeros(2 * (DIM + 2))
    check(ctx, ccall((:kamr_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int32, Ptr{Float64}), ctx.h, st.Δt, want, res))
    want && residual_comm_from!(ka, res)
    st.step += 1; st.sim_time += st.Δt                                 # counters of Theory/Iterate.jl:10-14
    return nothing
end

# the three public entry points one at a time (docs/src/methods_solve.md:23-26) — same results as step!
slope!(ctx::Context) = check(ctx, ccall((:kamr_slope, LIB), Cint, (Ptr{Cvoid},), ctx.h))
flux!(ctx::Context, ka::KA) = check(ctx, ccall((:kamr_flux, LIB), Cint, (Ptr{Cvoid}, Cdouble), ctx.h, ka.kinfo.status.Δt))
function iterate!(ctx::Context, ka::KA{DIM}) where {DIM}
    st = ka.kinfo.status
    want = st.step % ka.kinfo.config.solver.ST_CHECK_INTERVAL == 0
    res = zeros(2 * (DIM + 2))
    check(ctx, ccall((:kamr_iterate, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int32, Ptr{Float64}), ctx.h, st.Δt, want, res))
    want && residual_comm_from!(ka, res)
    st.step += 1; st.sim_time += st.Δt
    return nothing
end

"""
Before ps / vs adaptation, partition, save_result, save_for_restart: bring back what the event reads (SURVEY.md
Appendix D) and put it into `ka`.  `slopes = true` first runs `kamr_slope`, which also ships the mirrors' `sw` to the
ghosts as `sw_exchange!` does (the Löhner sensor reads them, Physical_space/Criteria.jl:103-169).
"""
function download!(ctx::Context, mask::UInt32 = DL_DF | DL_W | DL_PRIM | DL_QF; slopes::Bool = false)
    f = ctx.flat
    if slopes
        slope!(ctx); mask |= DL_SDF | DL_SW
    end
    GC.@preserve f check(ctx, ccall((:kamr_download_state, LIB), Cint,
        (Ptr{Cvoid}, UInt32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}), ctx.h, mask, f.df, f.sdf, f.flux, f.w, f.prim, f.qf, f.sw, f.mflux))
    scatter_state!(f, mask)
    return nothing
end

"""
    update_criterion!(ctx, ka)

Drop-in for `update_criterion!(ka)` (Physical_space/AMR.jl:256-286, including `apply_amr_buffer!` and its
`lohner_flag_exchange!`): the Löhner sensor of every local cell is evaluated on the device from the resident
`w` / `prim` / `sw` and only `PsData.lohner` ((DIM+2)×DIM per cell) comes back — no `df`, `sdf` or ghost data is
downloaded for a physical-space adaptation pass.  Runs `kamr_slope` first, as `ps_adaptive_mesh_refinement!` runs
`slope!` (AMR.jl:1109-1111).  `ps_refine_flag` / `ps_coarsen_flag` (Criteria.jl:69-101) then read `ps_data.lohner`
on the host unchanged.
"""
function update_criterion!(ctx::Context, ka::KA{DIM}) where {DIM}
    f = ctx.flat
    M = DIM + 2
    slope!(ctx)
    lohner = Vector{Float64}(undef, f.n_local * M * DIM)
    check(ctx, ccall((:kamr_ps_criterion, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Float64}, Ptr{Float64}),
                     ctx.h, ka.kinfo.config.solver.ADAPT_COEFFI_PS, lohner, C_NULL))
    for c in 1:f.n_local
        x = f.cells[c]
        x.bound_enc < 0 && continue                                   # solid cells keep theirs (AMR.jl:265)
        vec(x.lohner) .= @view lohner[(c-1)*M*DIM+1:c*M*DIM]           # column-major [(DIM+2) × DIM] == the ABI layout
    end
    return nothing
end

struct CVsAdapt                        # == kamr_vs_adapt (include/kamr.h)
    mode::Int32; maxlevel::Int32; trees::NTuple{3,Int32}; pad_::Int32
    vmin::NTuple{3,Float64}; vmax::NTuple{3,Float64}
    coeff_lohner::Float64; coeff_local::Float64; coeff_global::Float64
    vr_density::Float64; vr_energy::Float64
end

function CVsAdapt(ka::KA{DIM}, vr_density = 0.0, vr_energy = 0.0) where {DIM}
    cfg = ka.kinfo.config; s = cfg.solver; q = cfg.quadrature
    t3(f, fill) = ntuple(d -> d <= DIM ? f(d) : fill, 3)
    CVsAdapt(s.ADAPT_VS_MODE === :lohner ? 0 : 1, s.AMR_VS_MAXLEVEL, t3(d -> Int32(cfg.vs_trees_num[d]), Int32(1)), 0,
             t3(d -> Float64(q[2d-1]), 0.0), t3(d -> Float64(q[2d]), 1.0),
             s.ADAPT_COEFFI_VS_LOHNER, s.ADAPT_COEFFI_VS_LOCAL, s.ADAPT_COEFFI_VS_GLOBAL, vr_density, vr_energy)
end

"""
    vs_flags(ctx, ka) -> (refine_flags, coarsen_ok)

The per-velocity-point decisions of `vs_refine!` / `vs_coarsen!` (Velocity_space/AMR.jl:26-115) evaluated on the
device: one byte per point and decision comes back instead of `df` + `sdf`.  Both vectors are in the flat point order
(`flat.vs_off[c] + i`).  `vs_resolution` (AMR.jl:139-152) is evaluated on the device too; its two `MPI.Allreduce(MAX)`
stay here.  The host then runs `refine_grid_stream!` with `refine_flags`, and `coarsen_grid_stream!` with `coarsen_ok`
for the cells whose grid the refinement pass left alone; a cell whose grid changed re-evaluates `vs_coarsen!`'s loop
body on the host after fetching its `df` with `kamr_pack_cells` (the reference evaluates `coarsen_ok` on the refined
grid, whose `sdf` is zero, Rebuild.jl:81).
"""
function vs_flags(ctx::Context, ka::KA{DIM}) where {DIM}
    f = ctx.flat
    res = zeros(2)
    check(ctx, ccall((:kamr_vs_resolution, LIB), Cint, (Ptr{Cvoid}, Ref{CVsAdapt}, Ptr{Float64}), ctx.h, CVsAdapt(ka), res))
    res[1] = MPI.Allreduce(res[1], MPI.MAX, MPI.COMM_WORLD); res[2] = MPI.Allreduce(res[2], MPI.MAX, MPI.COMM_WORLD)
    npts = f.vs_off[f.n_local+1]
    refine = Vector{UInt8}(undef, npts); coarsen = Vector{UInt8}(undef, npts)
    check(ctx, ccall((:kamr_vs_criterion, LIB), Cint, (Ptr{Cvoid}, Ref{CVsAdapt}, Ptr{UInt8}, Ptr{UInt8}),
                     ctx.h, CVsAdapt(ka, res[1], res[2]), refine, coarsen))
    return refine, coarsen
end

"""
    vs_conserved_correction!(ctx, cells)

Drop-in for `vs_conserved_correction!` (Velocity_space/AMR.jl:120-133) after the re-flatten that follows a
velocity-space adaptation pass: `cells` are the flat ids (0-based) of the cells whose `va_flags` entry is set.
"""
function vs_conserved_correction!(ctx::Context, cells::Vector{Int32})
    check(ctx, ccall((:kamr_project_cells, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}), ctx.h, length(cells), cells))
    check(ctx, ccall((:kamr_exchange_df, LIB), Cint, (Ptr{Cvoid},), ctx.h))
    return nothing
end

"""
    migrate_begin!(ctx, dest_rank, src_rank, src_cells, src_points); reflatten!(ctx, p4est, ka); migrate_finish!(ctx, recv_cells)

`ps_partition!` (Parallel/Partition.jl) with the heavy payload moved device to device: `dest_rank[c]` is the new owner
of flat cell `c` (this rank included), known after `p4est_partition` from the new `global_first_quadrant`; `src_*` are
the `receive_nums` / `vs_nums` the reference exchanges first (Partition.jl:300-338) with this rank's own kept cells
listed too; `recv_cells` are the new flat ids in arrival order (ascending source rank, sender's order).  The host
still ships `bound_enc`, `solid_cell_index`, `vs_levels` and `vs_midpoints` as `transfer_wrap` does — only `w` and `df`
stay on the devices.
"""
function migrate_begin!(ctx::Context, dest_rank::Vector{Int32}, src_rank::Vector{Int32}, src_cells::Vector{Int32},
                        src_points::Vector{Int64})
    cells = collect(Int32(0):Int32(length(dest_rank) - 1))
    check(ctx, ccall((:kamr_migrate_begin, LIB), Cint,
                     (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}),
                     ctx.h, length(cells), cells, dest_rank, length(src_rank), src_rank, src_cells, src_points))
end
function migrate_finish!(ctx::Context, recv_cells::Vector{Int32})
    check(ctx, ccall((:kamr_migrate_finish, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}), ctx.h, length(recv_cells), recv_cells))
    check(ctx, ccall((:kamr_exchange_df, LIB), Cint, (Ptr{Cvoid},), ctx.h))
end

"""
The `solve!`-shaped loop with the device in it (compare Solver/Solver.jl:44-87).  The adapt events keep their cadence;
each one is bracketed by a download (what it reads) and a re-flatten (what it changed).
"""
function solve!(p4est, ka::KA; max_steps::Int = typemax(Int))
    ctx = create(ka)
    reflatten!(ctx, p4est, ka)
    s = ka.kinfo.config.solver
    try
        while !KitAMR.reached_max_time(ka) && ka.kinfo.status.step < max_steps
            if KitAMR.adapt_due(ka)                                    # the interval tests of adaptive_mesh_refinement!
                download!(ctx, DL_DF | DL_W | DL_PRIM; slopes = true)
                KitAMR.adaptive_mesh_refinement!(p4est, ka)            # includes amr_recover! and, when due, partition!
                reflatten!(ctx, p4est, ka)
            end
            KitAMR.limit_Δt!(ka)
            step!(ctx, ka)
            KitAMR.check!(p4est, ka)
            KitAMR.check_for_convergence(ka) && break
        end
        download!(ctx)
    finally
        destroy(ctx)
    end
    return nothing
end

"""
Partition weight for `partition!(p4est, weight)` (Parallel/Partition.jl:228): on the device a cell's cost is not
proportional to `vs_num` alone (DESIGN.md §6): relative to a regular cell, a cell with a neighbour on another velocity
grid costs ≈2.1×, a cell with hanging / domain / solid faces ≈3×, a solid ghost cell ≈4×, every solid face of a donor
adds ≈5.5×.
"""
function kamr_partition_weight(ps::AbstractPsData{DIM,NDF}) where {DIM,NDF}
    ps isa InsideSolidData && return Cint(0)
    n = ps.vs_data.vs_num
    ps.bound_enc < 0 && return Cint(4n)
    general = any(s -> s != 1, ps.neighbor.state[1:2*DIM])
    nsolid = count(d -> !isempty(d) && d[1] isa SolidNeighbor, ps.neighbor.data[1:2*DIM])
    mapped = any(d -> !isempty(d) && d[1] !== nothing && !(d[1] isa SolidNeighbor) && has_vs(d[1]) &&
                      d[1].vs_data.vs_num != n, ps.neighbor.data[1:2*DIM])
    return Cint(round(n * ((general ? 3.0 : mapped ? 2.1 : 1.0) + 5.5 * nsolid)))
end

# ---------------------------------------------------------------------------------------------------- the pin
"""
    dump_reference_step(p4est, ka, path; steps = 1)

Writes the fixture that pins the CPU oracle (and through it the device) against the REAL reference: the flat mesh and
state before, and `df`, `w`, `prim` after `steps` calls of the reference's own `slope!` / `flux!` / `iterate!`.  Plain
little-endian arrays with a text index (`tests/golden/read_reference_dump.py` reads it; no HDF5 needed on the
Python side).  Run it once for S0 (`test/runtests.jl`), `rp_2D.jl` and `cylinder.jl` and commit the result.
"""
function dump_reference_step(p4est, ka::KA, path::String; steps::Int = 1)
    flat = flatten(p4est, ka)
    mkpath(path)
    names = (:ds, :mid, :bound_enc, :ps_level, :cell_grid, :grid_off, :v_level, :v_weight, :v_mid, :nb_state, :nb_off,
             :nb_ids, :face_kind, :face_here, :face_there, :face_dir, :face_rot, :face_mid, :face_there_mid, :bc_type,
             :bc_prim, :peer_rank, :send_off, :send_cells, :recv_off, :solid_cell, :solid_nb_off, :solid_nb_ids, :sn_donor,
             :sn_solid, :sn_faceid, :sn_aux, :sn_normal, :sn_bc, :sn_nb_off, :sn_nb_ids, :cvc_off, :cvc_index, :cvc_gas_w,
             :cvc_solid_w, :df, :w, :prim)
    open(joinpath(path, "index.txt"), "w") do io
        println(io, "dim $(flat.dim)\nndf $(flat.ndf)\nn_local $(flat.n_local)\nn_ghost $(flat.n_ghost)\nn_solidnbr $(flat.n_solidnbr)")
        println(io, "ps_maxlevel $(flat.ps_maxlevel)\nps_minlevel $(flat.ps_minlevel)\ndt $(Float64(ka.kinfo.status.Δt))\nsteps $steps")
        g = ka.kinfo.config.gas; s = ka.kinfo.config.solver
        println(io, "K $(Float64(g.K))\nPr $(Float64(g.Pr))\ngamma $(Float64(g.γ))\nomega $(Float64(g.ω))\nmu_ref $(Float64(g.μᵣ))")
        println(io, "flux_type $(flux_code(s.flux))\nmarching $(march_code(s.time_marching))")
        for nm in names
            v = getfield(flat, nm)
            println(io, "array $nm $(eltype(v)) $(length(v))")
            write(joinpath(path, "$nm.bin"), v)
        end
    end
    for _ in 1:steps
        KitAMR.slope!(p4est, ka); KitAMR.flux!(p4est, ka); KitAMR.iterate!(p4est, ka)
    end
    gather_state!(flat)
    for nm in (:df, :w, :prim)
        write(joinpath(path, "after_$nm.bin"), getfield(flat, nm))
    end
    return path
end

end # module

/*
 * joltc_gpx.h — the subset of the joltc C API that NBT22/c-game-engine calls, served by libgpx (B200).
 *
 * SURVEY §8(b) shape B1: the engine and game sources keep calling `JPH_*`; they are compiled against this header
 * (the `include/joltc/...` files forward here, so `#include <joltc/joltc.h>` keeps working) and linked against
 * `libjoltc_gpx.so` instead of joltc + JoltPhysics.  Every declaration names the reference call site it serves
 * (paths relative to the reference tree).
 *
 * joltc itself (github.com/NBT22/joltc @ 226bd1e, a fork with `*_GAME` entry points and a `Transform` type) is NOT in
 * the reference tree — it is fetched at configure time (engine/CMakeLists.txt:53).  Type layouts and signatures below
 * are therefore INFERRED FROM THE CALL SITES: this is a source-level replacement (recompile the engine against it),
 * not a binary-compatible one.  Struct members the reference never touches are omitted.
 *
 * What the callbacks become on a GPU (device code cannot call back into the engine):
 *   - layer filters (pure functions of the layer in the reference) are evaluated once, at `*_Create`, into bit masks;
 *   - the laser BodyFilter is re-evaluated on the host for the live bodies before each filtered cast and pushed to the
 *     device as a per-body flag;
 *   - the character contact listener is fed from the device's added / persisted / removed list when
 *     JPH_PhysicsSystem_Update returns, on the calling thread, so handlers may create and destroy bodies.
 */
#ifndef JOLTC_GPX_H
#define JOLTC_GPX_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar types, constants, enums ------------------------------------------------------------------------------- */

typedef uint32_t JPH_BodyID;       /* stored in a LIST_UINT32, engine/src/structs/Map.c:52.  Low 23 bits: index (body slot, or
                                    * 0x400000 + k for static mesh bodies); top byte: sequence number, bumped when a slot is
                                    * reused, so calls with the id of a destroyed body find nothing (as in Jolt) */
typedef uint32_t JPH_SubShapeID;
typedef uint32_t JPH_ObjectLayer;  /* enum ObjectLayers, engine/include/engine/physics/Physics.h:36-42 */
typedef uint8_t JPH_BroadPhaseLayer;

#define JPH_BodyId_InvalidBodyID 0xFFFFFFFFu /* game/src/actor/prop/Door.c:163 */
#define JPH_BroadPhaseLayerInvalid ((JPH_BroadPhaseLayer)0xFF) /* engine/src/physics/Physics.c:31 */
#define JPH_DefaultConvexRadius 0.05f /* engine/src/assets/ModelLoader.c:332 */

typedef enum JPH_MotionType { JPH_MotionType_Static = 0, JPH_MotionType_Kinematic = 1, JPH_MotionType_Dynamic = 2 } JPH_MotionType;
typedef enum JPH_Activation { JPH_Activation_Activate = 0, JPH_Activation_DontActivate = 1 } JPH_Activation;
typedef enum JPH_PhysicsUpdateError /* engine/src/physics/MapPhysics.c:105-113 */
{
	JPH_PhysicsUpdateError_None = 0,
	JPH_PhysicsUpdateError_ManifoldCacheFull = 1,
	JPH_PhysicsUpdateError_BodyPairCacheFull = 2,
	JPH_PhysicsUpdateError_ContactConstraintsFull = 4,
} JPH_PhysicsUpdateError;
typedef enum JPH_AllowedDOFs /* game/src/actor/npc/TestActor.c:42-46 */
{
	JPH_AllowedDOFs_TranslationX = 1, JPH_AllowedDOFs_TranslationY = 2, JPH_AllowedDOFs_TranslationZ = 4,
	JPH_AllowedDOFs_RotationX = 8, JPH_AllowedDOFs_RotationY = 16, JPH_AllowedDOFs_RotationZ = 32,
	JPH_AllowedDOFs_All = 63,
} JPH_AllowedDOFs;
typedef enum JPH_OverrideMassProperties /* game/src/actor/prop/Physbox.c:31 */
{
	JPH_OverrideMassProperties_CalculateMassAndInertia = 0,
	JPH_OverrideMassProperties_CalculateInertia = 1,
	JPH_OverrideMassProperties_MassAndInertiaProvided = 2,
} JPH_OverrideMassProperties;
typedef enum JPH_GroundState /* engine/src/physics/PlayerPhysics.c:284 */
{
	JPH_GroundState_OnGround = 0,
	JPH_GroundState_OnSteepGround = 1,
	JPH_GroundState_NotSupported = 2,
	JPH_GroundState_InAir = 3,
} JPH_GroundState;

/* ---- math types (joltc/Math headers) ------------------------------------------------------------------------------------- */

typedef struct Vector3 { float x, y, z; } Vector3;
typedef Vector3 JPH_Vec3;
typedef Vector3 JPH_RVec3; /* single-precision build: used interchangeably with Vector3 (PlayerPhysics.c:106) */
typedef struct JPH_Quat { float x, y, z, w; } JPH_Quat;
typedef struct Transform { Vector3 position; JPH_Quat rotation; } Transform; /* joltc/Math/Transform.h (fork) */
typedef struct JPH_Mat44 { float m[16]; } JPH_Mat44; /* column-major, memcpy'd into a cglm mat4 (RenderingHelpers.c:110) */
typedef JPH_Mat44 JPH_RMat44;
typedef struct JPH_Plane { Vector3 normal; float distance; } JPH_Plane;
typedef struct JPH_Triangle { Vector3 v1, v2, v3; uint32_t materialIndex; } JPH_Triangle; /* MapLoader.c:231-247 */
typedef struct JPH_MassProperties { float mass; JPH_Mat44 inertia; } JPH_MassProperties;  /* Physbox.c:27-29 */
typedef struct JPH_RayCastResult /* PlayerPhysics.c:393-404, Laser.c:132 */
{
	JPH_BodyID bodyID;
	float fraction;
	JPH_SubShapeID subShapeID2; /* canonical face id here: triangle index in upload order / box face / 0 */
} JPH_RayCastResult;

extern const Vector3 Vector3_Zero, Vector3_AxisX, Vector3_AxisY, Vector3_AxisZ;
extern const JPH_Quat JPH_Quat_Identity;

/* Host math, no device work (SURVEY §8b "pure host math"). */
void Vector3_Add(const Vector3 *a, const Vector3 *b, Vector3 *out);                 /* Door.c:156 */
void Vector3_Subtract(const Vector3 *a, const Vector3 *b, Vector3 *out);            /* PlayerPhysics.c:351 */
void Vector3_MultiplyScalar(const Vector3 *a, float s, Vector3 *out);               /* PlayerPhysics.c:248 */
void Vector3_Normalized(const Vector3 *a, Vector3 *out);                            /* PlayerPhysics.c:244 */
float Vector3_Length(const Vector3 *a);
float Vector3_LengthSquared(const Vector3 *a);                                      /* PlayerPhysics.c:352 */
void JPH_Quat_Rotation(const Vector3 *axis, float angle, JPH_Quat *out);            /* PlayerPhysics.c:268 */
void JPH_Quat_Rotate(const JPH_Quat *q, const Vector3 *v, Vector3 *out);            /* PlayerPhysics.c:261 */
void JPH_Quat_RotateAxisZ(const JPH_Quat *q, Vector3 *out);                         /* Door.c:58, LaserEmitter.c:107 */
float JPH_Quat_GetRotationAngle(const JPH_Quat *q, const Vector3 *axis);            /* PlayerPhysics.c:269,510 */
void JPH_Quat_Normalized(const JPH_Quat *q, JPH_Quat *out);                         /* PlayerPhysics.c:375 */
void JPH_Quat_Multiply(const JPH_Quat *a, const JPH_Quat *b, JPH_Quat *out);        /* PlayerPhysics.c:515 */
void JPH_Quat_Lerp(const JPH_Quat *from, const JPH_Quat *to, float t, JPH_Quat *out); /* PlayerPhysics.c:374 */
void JPH_Quat_FromEulerAngles(const Vector3 *angles, JPH_Quat *out);                /* MapLoader.c:90 (X, then Y, then Z) */
void JPH_Quat_GetEulerAngles(const JPH_Quat *q, Vector3 *out);                      /* GlobalFog.c:45 */

/* ---- opaque objects ---------------------------------------------------------------------------------------------------- */

typedef struct JPH_JobSystem JPH_JobSystem;
typedef struct JPH_PhysicsSystem JPH_PhysicsSystem;
typedef struct JPH_BodyInterface JPH_BodyInterface;
typedef struct JPH_NarrowPhaseQuery JPH_NarrowPhaseQuery;
typedef struct JPH_Body JPH_Body;
typedef struct JPH_BodyCreationSettings JPH_BodyCreationSettings;
typedef struct JPH_Shape JPH_Shape;
typedef struct JPH_ShapeSettings JPH_ShapeSettings;
/* The reference casts between these freely ((JPH_Shape *)JPH_BoxShape_Create(..), (JPH_ShapeSettings *)compound..). */
typedef struct JPH_Shape JPH_BoxShape, JPH_ConvexHullShape, JPH_CylinderShape, JPH_CapsuleShape, JPH_MeshShape,
	JPH_StaticCompoundShape;
typedef struct JPH_ShapeSettings JPH_MeshShapeSettings, JPH_CompoundShapeSettings, JPH_StaticCompoundShapeSettings,
	JPH_EmptyShapeSettings;
typedef struct JPH_BroadPhaseLayerInterface JPH_BroadPhaseLayerInterface;
typedef struct JPH_ObjectLayerPairFilter JPH_ObjectLayerPairFilter;
typedef struct JPH_ObjectVsBroadPhaseLayerFilter JPH_ObjectVsBroadPhaseLayerFilter;
typedef struct JPH_BroadPhaseLayerFilter JPH_BroadPhaseLayerFilter;
typedef struct JPH_ObjectLayerFilter JPH_ObjectLayerFilter;
typedef struct JPH_BodyFilter JPH_BodyFilter;
typedef struct JPH_ShapeFilter JPH_ShapeFilter;
typedef struct JPH_CharacterVirtual JPH_CharacterVirtual;
typedef struct JPH_CharacterVirtual JPH_CharacterBase; /* (JPH_CharacterBase *)player->joltCharacter, PlayerPhysics.c:284 */
typedef struct JPH_CharacterContactListener JPH_CharacterContactListener;
typedef struct JPH_DebugRenderer JPH_DebugRenderer;
typedef struct JPH_BodyDrawFilter JPH_BodyDrawFilter;
typedef struct JPH_JobSystemThreadPoolConfig JPH_JobSystemThreadPoolConfig;

/* ---- lifecycle (engine/src/physics/Physics.c:72-87) ----------------------------------------------------------------- */

bool JPH_Init(void);     /* gpx_init on the device named by GPX_DEVICE (default 0) */
void JPH_Shutdown(void);
/* The tick runs on the GPU; the job system is a token the engine passes back into Update. */
JPH_JobSystem *JPH_JobSystemThreadPool_Create(const JPH_JobSystemThreadPoolConfig *config);
void JPH_JobSystem_Destroy(JPH_JobSystem *jobSystem);

/* ---- world + layer tables (Physics.c:20-105) -------------------------------------------------------------------------- */

typedef struct JPH_BroadPhaseLayerInterface_Impl { JPH_BroadPhaseLayer (*GetBroadPhaseLayer)(JPH_ObjectLayer layer); } JPH_BroadPhaseLayerInterface_Impl;
typedef struct JPH_ObjectLayerPairFilter_Impl { bool (*ShouldCollide)(JPH_ObjectLayer a, JPH_ObjectLayer b); } JPH_ObjectLayerPairFilter_Impl;
typedef struct JPH_ObjectVsBroadPhaseLayerFilter_Impl { bool (*ShouldCollide)(JPH_ObjectLayer a, JPH_BroadPhaseLayer b); } JPH_ObjectVsBroadPhaseLayerFilter_Impl;
/* Evaluated here, once, over the 4 object layers x 2 broad-phase layers the engine defines. */
JPH_BroadPhaseLayerInterface *JPH_BroadPhaseLayerInterface_Create(uint32_t numBroadPhaseLayers, const JPH_BroadPhaseLayerInterface_Impl *impl);
JPH_ObjectLayerPairFilter *JPH_ObjectLayerPairFilter_Create(const JPH_ObjectLayerPairFilter_Impl *impl);
JPH_ObjectVsBroadPhaseLayerFilter *JPH_ObjectVsBroadPhaseLayerFilter_Create(const JPH_ObjectVsBroadPhaseLayerFilter_Impl *impl);

typedef struct JPH_PhysicsSystemSettings /* Physics.c:91-97 */
{
	uint32_t maxBodies;             /* 0 = 64 (one ensemble-kernel world); the GPX_MAX_BODIES environment variable overrides */
	uint32_t numBodyMutexes;        /* unused */
	uint32_t maxBodyPairs;          /* unused */
	uint32_t maxContactConstraints; /* MAX_CONTACT_CONSTRAINTS: upper bound of the manifold pool */
	uint32_t _padding;
	JPH_BroadPhaseLayerInterface *broadPhaseLayerInterface;
	JPH_ObjectLayerPairFilter *objectLayerPairFilter;
	JPH_ObjectVsBroadPhaseLayerFilter *objectVsBroadPhaseLayerFilter;
} JPH_PhysicsSystemSettings;

/* The device collision matrix is the engine's (Physics.c:35-60).  Create returns NULL, with the reason on stderr, if the
 * callbacks describe a different one. */
JPH_PhysicsSystem *JPH_PhysicsSystem_Create(const JPH_PhysicsSystemSettings *settings);
void JPH_PhysicsSystem_Destroy(JPH_PhysicsSystem *system);                                  /* Physics.c:105 */
void JPH_PhysicsSystem_SetGravity(JPH_PhysicsSystem *system, const Vector3 *gravity);       /* Physics.c:99; before the first body */
JPH_BodyInterface *JPH_PhysicsSystem_GetBodyInterface(const JPH_PhysicsSystem *system);     /* 10 call sites */
const JPH_NarrowPhaseQuery *JPH_PhysicsSystem_GetNarrowPhaseQuery(const JPH_PhysicsSystem *system); /* PlayerPhysics.c:304 */
void JPH_PhysicsSystem_OptimizeBroadPhase(JPH_PhysicsSystem *system);                       /* MapLoader.c:273: LBVH build */
/* One fixed tick on the device (MapPhysics.c:105-108), then the transform mirror and the contact callbacks. */
JPH_PhysicsUpdateError JPH_PhysicsSystem_Update(JPH_PhysicsSystem *system, float deltaTime, int collisionSteps, JPH_JobSystem *jobSystem);

/* ---- shapes.  Ref-counted: constructors return +1, bodies take their own reference. ----------- */

JPH_BoxShape *JPH_BoxShape_Create(const Vector3 *halfExtent, float convexRadius);           /* ModelLoader.c:152, Trigger.c:37 */
JPH_ConvexHullShape *JPH_ConvexHullShape_Create(const Vector3 *points, uint32_t count, float maxConvexRadius); /* ModelLoader.c:330 */
JPH_CylinderShape *JPH_CylinderShape_Create(float halfHeight, float radius);                /* NpcJohn.c:29 */
JPH_CapsuleShape *JPH_CapsuleShape_Create(float halfHeightOfCylinder, float radius);        /* PlayerPhysics.c:176 (character only) */
JPH_MeshShapeSettings *JPH_MeshShapeSettings_Create(const JPH_Triangle *triangles, uint32_t count); /* ModelLoader.c:347 */
JPH_MeshShape *JPH_MeshShapeSettings_CreateShape(const JPH_MeshShapeSettings *settings);    /* ModelLoader.c:348 */
JPH_StaticCompoundShapeSettings *JPH_StaticCompoundShapeSettings_Create(void);              /* MapLoader.c:218 */
void JPH_CompoundShapeSettings_AddShape2(JPH_CompoundShapeSettings *settings, const Vector3 *position, const JPH_Quat *rotation,
										 const JPH_Shape *shape, uint32_t userData);           /* MapLoader.c:246 */
JPH_StaticCompoundShape *JPH_StaticCompoundShape_Create(const JPH_StaticCompoundShapeSettings *settings); /* MapLoader.c:255 */
JPH_EmptyShapeSettings *JPH_EmptyShapeSettings_Create(const Vector3 *centerOfMass);         /* Actor.c:153, Laser.c:114 */
void JPH_ShapeSettings_Destroy(JPH_ShapeSettings *settings);
void JPH_Shape_Destroy(JPH_Shape *shape);
/* Not joltc: 1 when the device primitive equals the shape (box, sphere-like hull, box-like hull, mesh, empty), 0 when a
 * bounding box stands in for it (cylinder, general hulls, several hulls in one compound). */
int JPH_GPX_ShapeIsExact(const JPH_Shape *shape);

/* ---- bodies (17 Create*Collider call sites, e.g. game/src/actor/prop/Physbox.c:19-38) --------------------------- */

JPH_BodyCreationSettings *JPH_BodyCreationSettings_Create2_GAME(const JPH_Shape *shape, const Transform *transform, JPH_MotionType motionType,
																JPH_ObjectLayer layer, void *userData);
JPH_BodyCreationSettings *JPH_BodyCreationSettings_Create_GAME(const JPH_ShapeSettings *shapeSettings, const Transform *transform,
															   JPH_MotionType motionType, JPH_ObjectLayer layer, void *userData); /* Actor.c:154 */
void JPH_BodyCreationSettings_Destroy(JPH_BodyCreationSettings *settings);
void JPH_BodyCreationSettings_SetFriction(JPH_BodyCreationSettings *settings, float friction);          /* MapLoader.c:263 */
void JPH_BodyCreationSettings_SetIsSensor(JPH_BodyCreationSettings *settings, bool isSensor);           /* Trigger.c:44 */
void JPH_BodyCreationSettings_SetMassPropertiesOverride(JPH_BodyCreationSettings *settings, const JPH_MassProperties *massProperties);
void JPH_BodyCreationSettings_SetOverrideMassProperties(JPH_BodyCreationSettings *settings, JPH_OverrideMassProperties mode);
void JPH_BodyCreationSettings_SetAllowedDOFs(JPH_BodyCreationSettings *settings, JPH_AllowedDOFs dofs); /* TestActor.c:42 */

/* Static bodies whose shape holds triangles join the static soup (ids >= 0x400000); everything else takes a body slot.
 * Returns JPH_BodyId_InvalidBodyID when the world is full or the shape cannot be represented (a mesh on a moving body). */
JPH_BodyID JPH_BodyInterface_CreateAndAddBody(JPH_BodyInterface *bi, const JPH_BodyCreationSettings *settings, JPH_Activation activation);
void JPH_BodyInterface_RemoveAndDestroyBody(JPH_BodyInterface *bi, JPH_BodyID body);                    /* Actor.c:68, Map.c:113 */
/* Getters read the host mirror written when Update returns (rows a9); they never wait for the device unless a create or
 * set happened since. */
void JPH_BodyInterface_GetPosition(JPH_BodyInterface *bi, JPH_BodyID body, JPH_RVec3 *out);             /* PlayerPhysics.c:347 */
void JPH_BodyInterface_GetRotation(JPH_BodyInterface *bi, JPH_BodyID body, JPH_Quat *out);              /* PlayerPhysics.c:367 */
void JPH_BodyInterface_GetPositionAndRotation(JPH_BodyInterface *bi, JPH_BodyID body, JPH_RVec3 *position, JPH_Quat *rotation); /* Camera.c:39 */
void JPH_BodyInterface_GetWorldTransform(JPH_BodyInterface *bi, JPH_BodyID body, JPH_RMat44 *out);      /* RenderingHelpers.c:110 */
uint64_t JPH_BodyInterface_GetUserData(JPH_BodyInterface *bi, JPH_BodyID body);                         /* PlayerPhysics.c:111 */
void JPH_BodyInterface_SetLinearVelocity(JPH_BodyInterface *bi, JPH_BodyID body, const Vector3 *velocity); /* Door.c:59 */
void JPH_BodyInterface_SetLinearAndAngularVelocity(JPH_BodyInterface *bi, JPH_BodyID body, const Vector3 *linear, const Vector3 *angular); /* PlayerPhysics.c:377 */
void JPH_BodyInterface_SetPosition(JPH_BodyInterface *bi, JPH_BodyID body, const JPH_RVec3 *position, JPH_Activation activation); /* Door.c:82 */
void JPH_BodyInterface_SetRotation(JPH_BodyInterface *bi, JPH_BodyID body, const JPH_Quat *rotation, JPH_Activation activation);  /* PlayerPhysics.c:381 */
uint64_t JPH_Body_GetUserData(const JPH_Body *body);            /* Laser.c:84 */
JPH_ObjectLayer JPH_Body_GetObjectLayer(const JPH_Body *body);  /* JoltDebugRenderer.c:16 */

/* ---- ray queries (PlayerPhysics.c:55-86,297-315; Laser.c:40-158) -------------------------------------------------- */

typedef struct JPH_BroadPhaseLayerFilter_Impl { bool (*ShouldCollide)(JPH_BroadPhaseLayer layer); } JPH_BroadPhaseLayerFilter_Impl;
typedef struct JPH_ObjectLayerFilter_Impl { bool (*ShouldCollide)(JPH_ObjectLayer layer); } JPH_ObjectLayerFilter_Impl;
typedef struct JPH_BodyFilter_Impl
{
	bool (*ShouldCollide)(JPH_BodyID body);
	bool (*ShouldCollideLocked)(const JPH_Body *body);
} JPH_BodyFilter_Impl;
typedef struct JPH_ShapeFilter_Impl { bool (*ShouldCollide)(const JPH_Shape *shape, const JPH_SubShapeID *subShapeID); } JPH_ShapeFilter_Impl;
JPH_BroadPhaseLayerFilter *JPH_BroadPhaseLayerFilter_Create(const JPH_BroadPhaseLayerFilter_Impl *impl);
void JPH_BroadPhaseLayerFilter_Destroy(JPH_BroadPhaseLayerFilter *filter);
JPH_ObjectLayerFilter *JPH_ObjectLayerFilter_Create(const JPH_ObjectLayerFilter_Impl *impl);
void JPH_ObjectLayerFilter_Destroy(JPH_ObjectLayerFilter *filter);
JPH_BodyFilter *JPH_BodyFilter_Create(const JPH_BodyFilter_Impl *impl);
void JPH_BodyFilter_Destroy(JPH_BodyFilter *filter);
JPH_ShapeFilter *JPH_ShapeFilter_Create(const JPH_ShapeFilter_Impl *impl); /* PlayerPhysics.c:160 (NULL impl) */
void JPH_ShapeFilter_Destroy(JPH_ShapeFilter *filter);

/* Closest hit along the transform's local -Z, `maxDistance` long; fraction in [0,1]. */
bool JPH_NarrowPhaseQuery_CastRay_GAME(const JPH_NarrowPhaseQuery *query, const Transform *origin, float maxDistance, JPH_RayCastResult *result,
									   const JPH_BroadPhaseLayerFilter *broadPhaseLayerFilter, const JPH_ObjectLayerFilter *objectLayerFilter);
/* The same from a body's transform, with the body filter; *hitPointOffset = hit point - ray origin. */
bool JPH_NarrowPhaseQuery_CastRay2_GAME(const JPH_NarrowPhaseQuery *query, JPH_BodyInterface *bi, JPH_BodyID originBody, float maxDistance,
										JPH_RayCastResult *result, Vector3 *hitPointOffset, const JPH_BroadPhaseLayerFilter *broadPhaseLayerFilter,
										const JPH_ObjectLayerFilter *objectLayerFilter, const JPH_BodyFilter *bodyFilter);

/* ---- player character (PlayerPhysics.c:89-194,283-294,439-453) ------------------------------------------------- */

typedef struct JPH_CharacterBaseSettings
{
	Vector3 up;
	JPH_Plane supportingVolume;
	float maxSlopeAngle; /* radians */
	bool enhancedInternalEdgeRemoval;
	const JPH_Shape *shape; /* the capsule */
} JPH_CharacterBaseSettings;
typedef struct JPH_CharacterVirtualSettings
{
	JPH_CharacterBaseSettings base;
	float mass;
	float maxStrength;
	Vector3 shapeOffset;
	float predictiveContactDistance;
	uint32_t maxCollisionIterations;
	uint32_t maxConstraintIterations;
	float minTimeRemaining;
	float collisionTolerance;
	float characterPadding;
	uint32_t maxNumHits;
	float hitReductionCosMaxAngle;
	float penetrationRecoverySpeed;
} JPH_CharacterVirtualSettings;
typedef struct JPH_ExtendedUpdateSettings /* PlayerPhysics.c:439-446 */
{
	Vector3 stickToFloorStepDown;
	Vector3 walkStairsStepUp;
	float walkStairsMinStepForward;
	float walkStairsStepForwardTest;
	float walkStairsCosAngleForwardContact;
	Vector3 walkStairsStepDownExtra;
} JPH_ExtendedUpdateSettings;
typedef struct JPH_CharacterContactSettings { bool canPushCharacter; bool canReceiveImpulses; } JPH_CharacterContactSettings;
typedef struct JPH_CharacterContactListener_Impl /* PlayerPhysics.c:146-151 */
{
	bool (*OnContactValidate)(const JPH_CharacterVirtual *character, JPH_BodyID body, JPH_SubShapeID subShape);
	void (*OnContactAdded)(const JPH_CharacterVirtual *character, JPH_BodyID body, JPH_SubShapeID subShape, const JPH_RVec3 *contactPosition,
						   const Vector3 *contactNormal, JPH_CharacterContactSettings *ioSettings);
	void (*OnContactPersisted)(const JPH_CharacterVirtual *character, JPH_BodyID body, JPH_SubShapeID subShape, const JPH_RVec3 *contactPosition,
							   const Vector3 *contactNormal, JPH_CharacterContactSettings *ioSettings);
	void (*OnContactRemoved)(const JPH_CharacterVirtual *character, JPH_BodyID body, JPH_SubShapeID subShape);
} JPH_CharacterContactListener_Impl;

/* Fills the members that are still zero with Jolt's defaults (the reference calls it AFTER its designated initialiser,
 * PlayerPhysics.c:177-185, so set members must survive). */
void JPH_CharacterVirtualSettings_Init(JPH_CharacterVirtualSettings *settings);
JPH_CharacterVirtual *JPH_CharacterVirtual_Create(const JPH_CharacterVirtualSettings *settings, const JPH_RVec3 *position, const JPH_Quat *rotation,
												  uint64_t userData, JPH_PhysicsSystem *system);
void JPH_CharacterVirtual_Destroy(JPH_CharacterVirtual *character);
void JPH_CharacterVirtual_SetUserData(JPH_CharacterVirtual *character, uint64_t userData);
uint64_t JPH_CharacterVirtual_GetUserData(const JPH_CharacterVirtual *character);
void JPH_CharacterVirtual_SetListener(JPH_CharacterVirtual *character, JPH_CharacterContactListener *listener);
void JPH_CharacterVirtual_SetPosition(JPH_CharacterVirtual *character, const JPH_RVec3 *position);
void JPH_CharacterVirtual_SetRotation(JPH_CharacterVirtual *character, const JPH_Quat *rotation); /* kept; a capsule about y is rotation-free */
void JPH_CharacterVirtual_GetPosition(const JPH_CharacterVirtual *character, JPH_RVec3 *out);    /* MapPhysics.c:77 */
void JPH_CharacterVirtual_GetLinearVelocity(const JPH_CharacterVirtual *character, Vector3 *out);
void JPH_CharacterVirtual_SetLinearVelocity(JPH_CharacterVirtual *character, const Vector3 *velocity);
JPH_GroundState JPH_CharacterBase_GetGroundState(const JPH_CharacterBase *character);
/* Moves the capsule on the device (collide and slide, ground state, stick to floor).  The listener's callbacks for the
 * contacts this produces are delivered when the following JPH_PhysicsSystem_Update returns; a contact whose
 * OnContactValidate answers false is not delivered (its push-out has already happened). */
void JPH_CharacterVirtual_ExtendedUpdate(JPH_CharacterVirtual *character, float deltaTime, const JPH_ExtendedUpdateSettings *settings,
										 JPH_ObjectLayer layer, const JPH_PhysicsSystem *system, const JPH_BodyFilter *bodyFilter,
										 const JPH_ShapeFilter *shapeFilter);
JPH_CharacterContactListener *JPH_CharacterContactListener_Create(const JPH_CharacterContactListener_Impl *impl);
void JPH_CharacterContactListener_Destroy(JPH_CharacterContactListener *listener);

/* ---- debug drawing (only with JPH_DEBUG_RENDERER, engine/src/debug/JoltDebugRenderer.c:35-52): accepted, draws nothing */

typedef struct JPH_DebugRenderer_Impl JPH_DebugRenderer_Impl;
typedef struct JPH_BodyDrawFilter_Impl { bool (*ShouldDraw)(void *userData, const JPH_Body *body); } JPH_BodyDrawFilter_Impl;
typedef struct JPH_DrawSettings JPH_DrawSettings;
JPH_DebugRenderer *JPH_DebugRenderer_Create(void *userData);
void JPH_DebugRenderer_Destroy(JPH_DebugRenderer *renderer);
void JPH_DebugRenderer_SetImpl(const JPH_DebugRenderer_Impl *impl);
JPH_BodyDrawFilter *JPH_BodyDrawFilter_Create(void *userData);
void JPH_BodyDrawFilter_Destroy(JPH_BodyDrawFilter *filter);
void JPH_BodyDrawFilter_SetImpl(const JPH_BodyDrawFilter_Impl *impl);
void JPH_PhysicsSystem_DrawBodies(const JPH_PhysicsSystem *system, const JPH_DrawSettings *settings, JPH_DebugRenderer *renderer,
								  const JPH_BodyDrawFilter *filter);

#ifdef __cplusplus
}
#endif
#endif /* JOLTC_GPX_H */

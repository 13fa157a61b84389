/*
 * engine_calls.c — a headless stand-in for the engine's physics thread, written against <joltc/...> exactly as the
 * reference's C sources are, and linked with libjoltc_gpx.so.  It walks the reference's own call sequences:
 *
 *   PhysicsInitGlobal / PhysicsInitMap      engine/src/physics/Physics.c:72-100
 *   CreatePlayerPhysics                     engine/src/physics/PlayerPhysics.c:173-194
 *   map collision load                      engine/src/assets/MapLoader.c:200-273
 *   CreateDynamicModelShape                 engine/src/assets/ModelLoader.c:324-343
 *   physbox / coin / door / laser colliders game/src/actor/prop/{Physbox,Coin,Door,Laser}.c
 *   MapFixedUpdate                          engine/src/physics/MapPhysics.c:58-119
 *
 * and prints what it observes (hex float bits) for tests/test_gpu_shim.py to compare with the oracle.
 *
 * usage: engine_calls <scene.bin> <ticks>
 * scene.bin: u32 nMeshes { f32 pos[3]; u32 nTris; f32 tris[nTris*9] } ; u32 nHullPoints; f32 pts[n*3] ;
 *            u32 nBoxes; f32 boxPos[n*3]
 */
#include <joltc/joltc.h>
#include <joltc/Math/Quat.h>
#include <joltc/Math/Transform.h>
#include <joltc/Math/Vector3.h>
#include <joltc/Physics/Body/BodyCreationSettings.h>
#include <joltc/Physics/Body/BodyInterface.h>
#include <joltc/Physics/Body/MassProperties.h>
#include <joltc/Physics/Collision/NarrowPhaseQuery.h>
#include <joltc/Physics/Collision/Shape/Shape.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum ObjectLayers { OBJECT_LAYER_STATIC, OBJECT_LAYER_DYNAMIC, OBJECT_LAYER_PLAYER, OBJECT_LAYER_SENSOR };
enum BroadPhaseLayers { BROAD_PHASE_LAYER_STATIC, BROAD_PHASE_LAYER_DYNAMIC, BROADPHASE_LAYER_MAX };
#define ACTOR_FLAG_CAN_BLOCK_LASERS 1u
#define GRAVITY (-9.81f)
#define PHYSICS_TARGET_TPS 60

typedef struct Actor
{
	const char *name;
	uint32_t flags;
	JPH_BodyID bodyId;
	JPH_BodyInterface *bodyInterface;
	int alive;
} Actor;

static JPH_PhysicsSystem *physicsSystem;
static JPH_JobSystem *jobSystem;
static JPH_CharacterVirtual *joltCharacter;
static int currentTick;

static uint32_t bits(float f)
{
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}

/* ---- layer tables, as Physics.c:20-70 ---------------------------------------------------------------------------------- */

static JPH_BroadPhaseLayer GetBroadPhaseLayer(const JPH_ObjectLayer inLayer)
{
	switch (inLayer)
	{
		case OBJECT_LAYER_STATIC:
		case OBJECT_LAYER_SENSOR: return BROAD_PHASE_LAYER_STATIC;
		case OBJECT_LAYER_DYNAMIC:
		case OBJECT_LAYER_PLAYER: return BROAD_PHASE_LAYER_DYNAMIC;
		default: return JPH_BroadPhaseLayerInvalid;
	}
}
static bool ObjectLayerShouldCollide(const JPH_ObjectLayer a, const JPH_ObjectLayer b)
{
	if (a == OBJECT_LAYER_DYNAMIC || a == OBJECT_LAYER_PLAYER)
		return b == OBJECT_LAYER_STATIC || b == OBJECT_LAYER_DYNAMIC || b == OBJECT_LAYER_SENSOR;
	return false;
}
static bool ObjectVsBroadPhaseLayerShouldCollide(const JPH_ObjectLayer layer, const JPH_BroadPhaseLayer bp)
{
	(void)bp;
	return GetBroadPhaseLayer(layer) != BROAD_PHASE_LAYER_STATIC;
}
static const JPH_ObjectVsBroadPhaseLayerFilter_Impl OBJECT_VS_BROAD_PHASE_LAYER_FILTER_IMPL = {.ShouldCollide = ObjectVsBroadPhaseLayerShouldCollide};
static const JPH_ObjectLayerPairFilter_Impl OBJECT_LAYER_PAIR_FILTER_IMPL = {.ShouldCollide = ObjectLayerShouldCollide};
static const JPH_BroadPhaseLayerInterface_Impl BROAD_PHASE_LAYER_INTERFACE_IMPL = {.GetBroadPhaseLayer = GetBroadPhaseLayer};

/* ---- ray filters, as PlayerPhysics.c:55-86 and Laser.c:40-101 ---------------------------------------------------------- */

static bool RaycastBroadPhaseLayerShouldCollide(const JPH_BroadPhaseLayer layer)
{
	return layer == BROAD_PHASE_LAYER_STATIC || layer == BROAD_PHASE_LAYER_DYNAMIC;
}
static bool RaycastObjectLayerShouldCollide(const JPH_ObjectLayer layer) { return layer == OBJECT_LAYER_STATIC || layer == OBJECT_LAYER_DYNAMIC; }
static bool TripleLaserBroadPhaseLayerShouldCollide(const JPH_BroadPhaseLayer layer) { return layer == BROAD_PHASE_LAYER_STATIC; }
static bool TripleLaserObjectLayerShouldCollide(const JPH_ObjectLayer layer) { return layer == OBJECT_LAYER_STATIC; }
static bool BodyFilterShouldCollide(const JPH_BodyID bodyId)
{
	JPH_BodyInterface *bodyInterface = JPH_PhysicsSystem_GetBodyInterface(physicsSystem);
	const Actor *actor = (const Actor *)JPH_BodyInterface_GetUserData(bodyInterface, bodyId);
	return !actor || ((actor->flags & ACTOR_FLAG_CAN_BLOCK_LASERS) == ACTOR_FLAG_CAN_BLOCK_LASERS);
}
static bool BodyFilterShouldCollideLocked(const JPH_Body *body)
{
	const Actor *actor = (const Actor *)JPH_Body_GetUserData(body);
	return !actor || ((actor->flags & ACTOR_FLAG_CAN_BLOCK_LASERS) == ACTOR_FLAG_CAN_BLOCK_LASERS);
}
static const JPH_BroadPhaseLayerFilter_Impl RAYCAST_BP_IMPL = {.ShouldCollide = RaycastBroadPhaseLayerShouldCollide};
static const JPH_ObjectLayerFilter_Impl RAYCAST_OL_IMPL = {.ShouldCollide = RaycastObjectLayerShouldCollide};
static const JPH_BroadPhaseLayerFilter_Impl TRIPLE_BP_IMPL = {.ShouldCollide = TripleLaserBroadPhaseLayerShouldCollide};
static const JPH_ObjectLayerFilter_Impl TRIPLE_OL_IMPL = {.ShouldCollide = TripleLaserObjectLayerShouldCollide};
static const JPH_BodyFilter_Impl BODY_FILTER_IMPL = {.ShouldCollide = BodyFilterShouldCollide, .ShouldCollideLocked = BodyFilterShouldCollideLocked};

/* ---- character listener, as PlayerPhysics.c:89-152 --------------------------------------------------------------------- */

static Actor *coinActor;

static bool OnContactValidate(const JPH_CharacterVirtual *character, const JPH_BodyID bodyId, JPH_SubShapeID subShapeId)
{
	(void)subShapeId;
	(void)bodyId;
	return JPH_CharacterVirtual_GetUserData(character) == 0xC0FFEEull;
}
static void OnContactAdded(const JPH_CharacterVirtual *character, const JPH_BodyID bodyId, JPH_SubShapeID subShapeId,
						   const JPH_RVec3 *contactPosition, const Vector3 *contactNormal, JPH_CharacterContactSettings *ioSettings)
{
	(void)character; (void)subShapeId; (void)contactPosition; (void)contactNormal;
	JPH_BodyInterface *bodyInterface = JPH_PhysicsSystem_GetBodyInterface(physicsSystem);
	Actor *actor = (Actor *)JPH_BodyInterface_GetUserData(bodyInterface, bodyId);
	printf("E %d added %08x %s\n", currentTick, bodyId, actor ? actor->name : "-");
	if (actor)
	{
		ioSettings->canPushCharacter = false;
		if (actor == coinActor && actor->alive)
		{
			/* CoinOnPlayerContactAdded -> RemoveActor -> ActorDestroy (Coin.c:85, Actor.c:68) */
			JPH_BodyInterface_RemoveAndDestroyBody(actor->bodyInterface, actor->bodyId);
			actor->alive = 0;
		}
	}
}
static void OnContactPersisted(const JPH_CharacterVirtual *character, const JPH_BodyID bodyId, JPH_SubShapeID subShapeId,
							   const JPH_RVec3 *contactPosition, const Vector3 *contactNormal, JPH_CharacterContactSettings *ioSettings)
{
	(void)character; (void)subShapeId; (void)contactPosition; (void)contactNormal; (void)ioSettings;
	printf("E %d persisted %08x\n", currentTick, bodyId);
}
static void OnContactRemoved(const JPH_CharacterVirtual *character, const JPH_BodyID bodyId, JPH_SubShapeID subShapeId)
{
	(void)character; (void)subShapeId;
	printf("E %d removed %08x\n", currentTick, bodyId);
}
static const JPH_CharacterContactListener_Impl CONTACT_LISTENER_IMPL = {
	.OnContactValidate = OnContactValidate,
	.OnContactAdded = OnContactAdded,
	.OnContactPersisted = OnContactPersisted,
	.OnContactRemoved = OnContactRemoved,
};

/* ---- scene file ----------------------------------------------------------------------------------------------------------- */

static uint32_t ReadU32(FILE *f)
{
	uint32_t v = 0;
	if (fread(&v, 4, 1, f) != 1) exit(3);
	return v;
}
static void ReadFloats(FILE *f, float *out, size_t n)
{
	if (n && fread(out, 4, n, f) != n) exit(3);
}

static void PrintBody(const char *tag, JPH_BodyInterface *bodyInterface, const JPH_BodyID id)
{
	Vector3 p = {0};
	JPH_Quat q = {0};
	JPH_BodyInterface_GetPositionAndRotation(bodyInterface, id, &p, &q);
	printf("X %d %s %08x %08x %08x %08x %08x %08x %08x %08x\n", currentTick, tag, id, bits(p.x), bits(p.y), bits(p.z), bits(q.x), bits(q.y),
		   bits(q.z), bits(q.w));
}

int main(int argc, char **argv)
{
	if (argc < 3) return 2;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 2;
	const int ticks = atoi(argv[2]);

	/* PhysicsInitGlobal */
	if (!JPH_Init()) return 4;
	jobSystem = JPH_JobSystemThreadPool_Create(NULL);
	JPH_CharacterContactListener *contactListener = JPH_CharacterContactListener_Create(&CONTACT_LISTENER_IMPL);
	JPH_ShapeFilter *shapeFilter = JPH_ShapeFilter_Create(NULL);
	JPH_BroadPhaseLayerFilter *actorRaycastBroadPhaseLayerFilter = JPH_BroadPhaseLayerFilter_Create(&RAYCAST_BP_IMPL);
	JPH_ObjectLayerFilter *actorRaycastObjectLayerFilter = JPH_ObjectLayerFilter_Create(&RAYCAST_OL_IMPL);
	JPH_BroadPhaseLayerFilter *tripleLaserBroadPhaseLayerFilter = JPH_BroadPhaseLayerFilter_Create(&TRIPLE_BP_IMPL);
	JPH_ObjectLayerFilter *tripleLaserObjectLayerFilter = JPH_ObjectLayerFilter_Create(&TRIPLE_OL_IMPL);
	JPH_BodyFilter *bodyFilter = JPH_BodyFilter_Create(&BODY_FILTER_IMPL);

	/* PhysicsInitMap */
	const JPH_PhysicsSystemSettings physicsSystemSettings = {
		.maxContactConstraints = 16384,
		.broadPhaseLayerInterface = JPH_BroadPhaseLayerInterface_Create(BROADPHASE_LAYER_MAX, &BROAD_PHASE_LAYER_INTERFACE_IMPL),
		.objectLayerPairFilter = JPH_ObjectLayerPairFilter_Create(&OBJECT_LAYER_PAIR_FILTER_IMPL),
		.objectVsBroadPhaseLayerFilter = JPH_ObjectVsBroadPhaseLayerFilter_Create(&OBJECT_VS_BROAD_PHASE_LAYER_FILTER_IMPL),
	};
	physicsSystem = JPH_PhysicsSystem_Create(&physicsSystemSettings);
	if (!physicsSystem) return 5;
	JPH_PhysicsSystem_SetGravity(physicsSystem, &(Vector3){0, GRAVITY, 0});
	JPH_BodyInterface *bodyInterface = JPH_PhysicsSystem_GetBodyInterface(physicsSystem);

	/* CreatePlayerPhysics */
	Transform playerTransform = {.position = {-0.2f, 0.0f, -0.5f}, .rotation = JPH_Quat_Identity};
	{
		JPH_Shape *shape = (JPH_Shape *)JPH_CapsuleShape_Create(0.2f, 0.25f);
		JPH_CharacterVirtualSettings characterSettings = {
			.base.supportingVolume.normal = Vector3_AxisY,
			.base.supportingVolume.distance = 0.25f,
			.base.maxSlopeAngle = 50.0f * 3.14159265358979323846f / 180.0f,
			.base.enhancedInternalEdgeRemoval = true,
			.base.shape = shape,
			.mass = 10.0f,
		};
		JPH_CharacterVirtualSettings_Init(&characterSettings);
		joltCharacter = JPH_CharacterVirtual_Create(&characterSettings, &playerTransform.position, NULL, 0, physicsSystem);
		if (!joltCharacter) return 6;
		JPH_CharacterVirtual_SetUserData(joltCharacter, 0xC0FFEEull);
		JPH_CharacterVirtual_SetListener(joltCharacter, contactListener);
		JPH_Shape_Destroy(shape);
	}

	/* map collision meshes */
	const uint32_t numCollisionMeshes = ReadU32(f);
	JPH_BodyID firstMapBody = JPH_BodyId_InvalidBodyID;
	for (uint32_t i = 0; i < numCollisionMeshes; i++)
	{
		Transform collisionXfm = {.rotation = JPH_Quat_Identity};
		ReadFloats(f, &collisionXfm.position.x, 3);
		const uint32_t numTriangles = ReadU32(f);
		JPH_Triangle *tris = malloc(sizeof(JPH_Triangle) * numTriangles);
		for (uint32_t k = 0; k < numTriangles; k++)
		{
			tris[k].materialIndex = 0;
			ReadFloats(f, &tris[k].v1.x, 3);
			ReadFloats(f, &tris[k].v2.x, 3);
			ReadFloats(f, &tris[k].v3.x, 3);
		}
		JPH_StaticCompoundShapeSettings *compoundShapeSettings = JPH_StaticCompoundShapeSettings_Create();
		JPH_MeshShapeSettings *settings = JPH_MeshShapeSettings_Create(tris, numTriangles);
		JPH_Shape *subShape = (JPH_Shape *)JPH_MeshShapeSettings_CreateShape(settings);
		JPH_ShapeSettings_Destroy((JPH_ShapeSettings *)settings);
		JPH_CompoundShapeSettings_AddShape2((JPH_CompoundShapeSettings *)compoundShapeSettings, &Vector3_Zero, &JPH_Quat_Identity, subShape, 0);
		JPH_Shape_Destroy(subShape);
		free(tris);
		JPH_Shape *shape = (JPH_Shape *)JPH_StaticCompoundShape_Create(compoundShapeSettings);
		JPH_BodyCreationSettings *bodyCreationSettings = JPH_BodyCreationSettings_Create2_GAME(shape, &collisionXfm, JPH_MotionType_Static,
																							   OBJECT_LAYER_STATIC, 0);
		JPH_BodyCreationSettings_SetFriction(bodyCreationSettings, 4.25f);
		const JPH_BodyID body = JPH_BodyInterface_CreateAndAddBody(bodyInterface, bodyCreationSettings, JPH_Activation_Activate);
		if (i == 0) firstMapBody = body;
		JPH_BodyCreationSettings_Destroy(bodyCreationSettings);
		JPH_ShapeSettings_Destroy((JPH_ShapeSettings *)compoundShapeSettings);
		JPH_Shape_Destroy(shape);
	}
	JPH_PhysicsSystem_OptimizeBroadPhase(physicsSystem);
	printf("M %u %08x\n", numCollisionMeshes, firstMapBody);

	/* CreateDynamicModelShape: one hull in a static compound */
	const uint32_t numPoints = ReadU32(f);
	Vector3 *points = malloc(sizeof(Vector3) * numPoints);
	ReadFloats(f, &points[0].x, 3ull * numPoints);
	JPH_Shape *collisionModelShape;
	{
		JPH_StaticCompoundShapeSettings *compoundShapeSettings = JPH_StaticCompoundShapeSettings_Create();
		JPH_Shape *hullShape = (JPH_Shape *)JPH_ConvexHullShape_Create(points, numPoints, JPH_DefaultConvexRadius);
		JPH_CompoundShapeSettings_AddShape2((JPH_CompoundShapeSettings *)compoundShapeSettings, &Vector3_Zero, &JPH_Quat_Identity, hullShape, 0);
		JPH_Shape_Destroy(hullShape);
		collisionModelShape = (JPH_Shape *)JPH_StaticCompoundShape_Create(compoundShapeSettings);
		JPH_ShapeSettings_Destroy((JPH_ShapeSettings *)compoundShapeSettings);
	}
	free(points);
	printf("S %d\n", JPH_GPX_ShapeIsExact(collisionModelShape));

	/* actors: physboxes */
	const uint32_t numBoxes = ReadU32(f);
	Actor *boxes = calloc(numBoxes, sizeof(Actor));
	for (uint32_t i = 0; i < numBoxes; i++)
	{
		Transform transform = {.rotation = JPH_Quat_Identity};
		ReadFloats(f, &transform.position.x, 3);
		Actor *this = &boxes[i];
		this->name = "prop_physbox";
		this->flags = ACTOR_FLAG_CAN_BLOCK_LASERS;
		this->bodyInterface = bodyInterface;
		this->alive = 1;
		JPH_BodyCreationSettings *bodyCreationSettings = JPH_BodyCreationSettings_Create2_GAME(collisionModelShape, &transform, JPH_MotionType_Dynamic,
																							   OBJECT_LAYER_DYNAMIC, this);
		const JPH_MassProperties massProperties = {.mass = 10.0f};
		JPH_BodyCreationSettings_SetMassPropertiesOverride(bodyCreationSettings, &massProperties);
		JPH_BodyCreationSettings_SetOverrideMassProperties(bodyCreationSettings, JPH_OverrideMassProperties_CalculateInertia);
		this->bodyId = JPH_BodyInterface_CreateAndAddBody(this->bodyInterface, bodyCreationSettings, JPH_Activation_Activate);
		JPH_BodyCreationSettings_Destroy(bodyCreationSettings);
	}
	fclose(f);

	/* a coin: sensor box on the SENSOR layer (Coin.c:40-55) */
	Actor coin = {.name = "prop_coin", .flags = 0, .bodyInterface = bodyInterface, .alive = 1};
	coinActor = &coin;
	{
		const Transform transform = {.position = {-1.0f, -1.25f, -0.5f}, .rotation = JPH_Quat_Identity};
		JPH_Shape *shape = (JPH_Shape *)JPH_BoxShape_Create((Vector3[]){{0.25f, 0.25f, 0.25f}}, JPH_DefaultConvexRadius);
		JPH_BodyCreationSettings *bodyCreationSettings = JPH_BodyCreationSettings_Create2_GAME(shape, &transform, JPH_MotionType_Static,
																							   OBJECT_LAYER_SENSOR, &coin);
		JPH_BodyCreationSettings_SetIsSensor(bodyCreationSettings, true);
		coin.bodyId = JPH_BodyInterface_CreateAndAddBody(bodyInterface, bodyCreationSettings, JPH_Activation_Activate);
		JPH_Shape_Destroy(shape);
		JPH_BodyCreationSettings_Destroy(bodyCreationSettings);
	}

	/* a door: kinematic flat hull on the STATIC layer (Door.c:107-126, ActorWall.c:20-49) */
	Actor door = {.name = "prop_door", .flags = ACTOR_FLAG_CAN_BLOCK_LASERS, .bodyInterface = bodyInterface, .alive = 1};
	const Vector3 doorClosed = {1.5f, -1.0f, -1.5f};
	{
		const Vector3 wallPoints[4] = {{0, -0.5f, -0.5f}, {0, -0.5f, 0.5f}, {0, 0.5f, -0.5f}, {0, 0.5f, 0.5f}};
		JPH_Shape *shape = (JPH_Shape *)JPH_ConvexHullShape_Create(wallPoints, 4, JPH_DefaultConvexRadius);
		const Transform transform = {.position = doorClosed, .rotation = JPH_Quat_Identity};
		JPH_BodyCreationSettings *bodyCreationSettings = JPH_BodyCreationSettings_Create2_GAME(shape, &transform, JPH_MotionType_Kinematic,
																							   OBJECT_LAYER_STATIC, &door);
		const JPH_MassProperties massProperties = {.mass = 1.0f};
		JPH_BodyCreationSettings_SetMassPropertiesOverride(bodyCreationSettings, &massProperties);
		JPH_BodyCreationSettings_SetOverrideMassProperties(bodyCreationSettings, JPH_OverrideMassProperties_CalculateInertia);
		door.bodyId = JPH_BodyInterface_CreateAndAddBody(bodyInterface, bodyCreationSettings, JPH_Activation_Activate);
		printf("S %d\n", JPH_GPX_ShapeIsExact(shape));
		JPH_Shape_Destroy(shape);
		JPH_BodyCreationSettings_Destroy(bodyCreationSettings);
	}

	/* two lasers: empty bodies (Laser.c:104-124); the second one is the 'triple' kind that only sees map geometry */
	Actor lasers[2] = {{.name = "laser", .bodyInterface = bodyInterface, .alive = 1}, {.name = "laser3", .bodyInterface = bodyInterface, .alive = 1}};
	for (int i = 0; i < 2; i++)
	{
		/* looking down -Z from behind the column at box height / looking along +Z (rotation pi about Y) */
		const Transform transform = {.position = {0.0f, i == 0 ? -1.25f : -1.0f, i == 0 ? 1.0f : -3.0f},
									 .rotation = i == 0 ? JPH_Quat_Identity : (JPH_Quat){0.0f, 1.0f, 0.0f, 0.0f}};
		JPH_ShapeSettings *shapeSettings = (JPH_ShapeSettings *)JPH_EmptyShapeSettings_Create(&Vector3_Zero);
		JPH_BodyCreationSettings *bodyCreationSettings = JPH_BodyCreationSettings_Create_GAME(shapeSettings, &transform, JPH_MotionType_Static,
																							  OBJECT_LAYER_STATIC, &lasers[i]);
		lasers[i].bodyId = JPH_BodyInterface_CreateAndAddBody(bodyInterface, bodyCreationSettings, JPH_Activation_DontActivate);
		JPH_ShapeSettings_Destroy(shapeSettings);
		JPH_BodyCreationSettings_Destroy(bodyCreationSettings);
	}
	printf("I coin %08x door %08x laser %08x laser3 %08x\n", coin.bodyId, door.bodyId, lasers[0].bodyId, lasers[1].bodyId);

	/* MapFixedUpdate x ticks */
	const double delta = 1.0;
	for (currentTick = 1; currentTick <= ticks; currentTick++)
	{
		/* MovePlayer: stand for 40 ticks, walk towards -x (the coin), later towards +x */
		Vector3 moveVec = Vector3_Zero;
		if (currentTick > 40) moveVec.x = currentTick <= 140 ? -1.5f : 1.5f;
		if (JPH_CharacterBase_GetGroundState((JPH_CharacterBase *)joltCharacter) != JPH_GroundState_OnGround)
		{
			Vector3 oldVelocity;
			JPH_CharacterVirtual_GetLinearVelocity(joltCharacter, &oldVelocity);
			moveVec.y += oldVelocity.y + (float)(GRAVITY * (delta / PHYSICS_TARGET_TPS));
		}
		JPH_CharacterVirtual_SetLinearVelocity(joltCharacter, &moveVec);

		const float deltaTime = (float)delta / PHYSICS_TARGET_TPS;

		/* UpdatePlayer: the targeting ray from the camera, then the character */
		{
			const Transform camera = {.position = {0.0f, -1.25f, 2.0f}, .rotation = JPH_Quat_Identity};
			JPH_RayCastResult raycastResult = {0};
			const JPH_NarrowPhaseQuery *narrowPhaseQuery = JPH_PhysicsSystem_GetNarrowPhaseQuery(physicsSystem);
			const bool hit = JPH_NarrowPhaseQuery_CastRay_GAME(narrowPhaseQuery, &camera, 10.0f, &raycastResult, actorRaycastBroadPhaseLayerFilter,
															   actorRaycastObjectLayerFilter);
			if (currentTick % 20 == 1)
				printf("R %d camera %d %08x %08x %08x\n", currentTick, hit, raycastResult.bodyID, bits(raycastResult.fraction), raycastResult.subShapeID2);
		}
		const JPH_ExtendedUpdateSettings extendedUpdateSettings = {
			.stickToFloorStepDown.y = -0.25f,
			.walkStairsStepUp.y = 0.25f,
			.walkStairsMinStepForward = 0.02f,
			.walkStairsStepForwardTest = 0.15f,
			.walkStairsCosAngleForwardContact = cosf(75.0f * 3.14159265358979323846f / 180.0f),
			.walkStairsStepDownExtra = Vector3_Zero,
		};
		JPH_CharacterVirtual_ExtendedUpdate(joltCharacter, deltaTime, &extendedUpdateSettings, OBJECT_LAYER_PLAYER, physicsSystem, NULL, shapeFilter);
		JPH_CharacterVirtual_GetPosition(joltCharacter, &playerTransform.position);

		/* actor updates: the door opens between ticks 60 and 120 (Door.c:53-105), the lasers cast (Laser.c:127-158) */
		if (currentTick == 60) JPH_BodyInterface_SetLinearVelocity(door.bodyInterface, door.bodyId, &(Vector3){0.0f, 0.0f, 1.0f});
		if (currentTick == 120)
		{
			JPH_BodyInterface_SetLinearVelocity(door.bodyInterface, door.bodyId, &Vector3_Zero);
			JPH_BodyInterface_SetPosition(door.bodyInterface, door.bodyId, &(Vector3){doorClosed.x, doorClosed.y, doorClosed.z + 1.0f},
										  JPH_Activation_DontActivate);
		}
		if (currentTick == 200) boxes[0].flags = 0; /* the bottom box stops blocking lasers */
		for (int i = 0; i < 2; i++)
		{
			JPH_RayCastResult result = {0};
			Vector3 hitPointOffset = {0};
			const bool hit = JPH_NarrowPhaseQuery_CastRay2_GAME(JPH_PhysicsSystem_GetNarrowPhaseQuery(physicsSystem), lasers[i].bodyInterface,
																lasers[i].bodyId, 50.0f, &result, &hitPointOffset,
																i == 1 ? tripleLaserBroadPhaseLayerFilter : actorRaycastBroadPhaseLayerFilter,
																i == 1 ? tripleLaserObjectLayerFilter : actorRaycastObjectLayerFilter, bodyFilter);
			if (currentTick % 20 == 1 || currentTick == 200)
				printf("R %d %s %d %08x %08x %08x %08x\n", currentTick, lasers[i].name, hit, result.bodyID, bits(result.fraction), result.subShapeID2,
					   bits(hitPointOffset.z));
		}

		const JPH_PhysicsUpdateError result = JPH_PhysicsSystem_Update(physicsSystem, deltaTime, 2, jobSystem);
		if (result != JPH_PhysicsUpdateError_None)
		{
			printf("F %d %d\n", currentTick, (int)result);
			return 7;
		}

		if (currentTick % 20 == 0 || currentTick == ticks)
		{
			for (uint32_t i = 0; i < numBoxes; i++) PrintBody("box", bodyInterface, boxes[i].bodyId);
			PrintBody("door", bodyInterface, door.bodyId);
			Vector3 v = {0};
			JPH_CharacterVirtual_GetLinearVelocity(joltCharacter, &v);
			printf("C %d %08x %08x %08x %08x %08x %08x %d\n", currentTick, bits(playerTransform.position.x), bits(playerTransform.position.y),
				   bits(playerTransform.position.z), bits(v.x), bits(v.y), bits(v.z),
				   (int)JPH_CharacterBase_GetGroundState((JPH_CharacterBase *)joltCharacter));
		}
	}
	JPH_RMat44 matrix;
	JPH_BodyInterface_GetWorldTransform(bodyInterface, boxes[0].bodyId, &matrix);
	printf("W %08x %08x %08x %08x\n", bits(matrix.m[12]), bits(matrix.m[13]), bits(matrix.m[14]), bits(matrix.m[15]));

	/* teardown in the engine's order: actors, map bodies, character, system, globals */
	for (uint32_t i = 0; i < numBoxes; i++) JPH_BodyInterface_RemoveAndDestroyBody(bodyInterface, boxes[i].bodyId);
	JPH_BodyInterface_RemoveAndDestroyBody(bodyInterface, firstMapBody);
	JPH_Shape_Destroy(collisionModelShape);
	JPH_CharacterVirtual_Destroy(joltCharacter);
	JPH_PhysicsSystem_Destroy(physicsSystem);
	JPH_CharacterContactListener_Destroy(contactListener);
	JPH_ShapeFilter_Destroy(shapeFilter);
	JPH_BroadPhaseLayerFilter_Destroy(actorRaycastBroadPhaseLayerFilter);
	JPH_ObjectLayerFilter_Destroy(actorRaycastObjectLayerFilter);
	JPH_BroadPhaseLayerFilter_Destroy(tripleLaserBroadPhaseLayerFilter);
	JPH_ObjectLayerFilter_Destroy(tripleLaserObjectLayerFilter);
	JPH_BodyFilter_Destroy(bodyFilter);
	JPH_JobSystem_Destroy(jobSystem);
	JPH_Shutdown();
	free(boxes);
	printf("done\n");
	return 0;
}
